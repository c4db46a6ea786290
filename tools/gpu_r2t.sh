#!/bin/bash
# shared initial fields for the warp-per-trajectory (random-site) kernel: parity, memcheck, bench sub-record
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py tests/test_harness.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=300 2>&1 | tail -6
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout=150 -k "generic_kernel or config1 or tiny_and_block" 2>&1 | tail -4
timeout 300 python bench.py --workload random_site --steps 2 --warmup 1 > gpurun_out/bench_random_$TAG.json 2> gpurun_out/bench_random_$TAG.err
python - $TAG <<'PY'
import json,sys
d=json.loads([l for l in open("gpurun_out/bench_random_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print(d["metric"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["achieved"], d["e2e"]["value"], d["gpu_launches"])
PY
timeout 200 python tools/random_mode_probe.py 2>&1 | cut -c1-260
echo "== U=16 variant (build/ab/genu16)"
OSA_LIB_PATH=build/ab/genu16/libonesolver_b200.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout=150 -k "generic_kernel or config1" 2>&1 | tail -2
OSA_LIB_PATH=build/ab/genu16/libonesolver_b200.so timeout 300 python bench.py --workload random_site --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(d['metric'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['achieved'])"
