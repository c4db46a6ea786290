#!/bin/bash
# round-2 call N: PT / PA with carried fields: parity + probe (tight timeouts)
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pa.py tests/test_gpu_pt.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=120 2>&1 | tail -12
timeout 200 python tools/pa_probe.py > gpurun_out/pa_probe_$TAG.log 2>&1; cut -c1-330 gpurun_out/pa_probe_$TAG.log
timeout 120 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_pa.py -m gpu -q -x -p no:cacheprovider --timeout=100 -k "matches or offset" 2>&1 | tail -4
