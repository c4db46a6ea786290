#!/bin/bash
# round-2 call J: population-annealing tests after the device-side fill, sampler probe
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pa.py tests/test_gpu_pt.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=600 2>&1 | tail -6
timeout 600 python tools/pa_probe.py > gpurun_out/pa_probe_$TAG.log 2>&1; cat gpurun_out/pa_probe_$TAG.log | cut -c1-330
