#!/bin/bash
# round-2 call H: full GPU suite, memcheck of the new kernels (sparse grouped layout, population annealing), bench
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider --tb=short --timeout=900 2>&1 | tail -30 > gpurun_out/pytest_$TAG.log
tail -6 gpurun_out/pytest_$TAG.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_pa.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout=1100 \
  -k "(sparse and not tiny) or pa_matches or pa_population" > gpurun_out/memcheck_sparse_pa_$TAG.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/memcheck_sparse_pa_$TAG.log | cut -c1-200
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python - $TAG <<'PY'
import json, sys
d=json.loads([l for l in open("gpurun_out/bench_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["nccl"])
for o in d["other_configs"]:
    print(o["metric"], o["value"], o["ms_per_step"], o["roofline"]["frac"], o["e2e"]["value"])
PY
