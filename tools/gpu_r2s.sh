#!/bin/bash
# sparse kernel with branch-free bookkeeping: parity (all sparse tests + fuzz), config 4 bench line
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_sparse_fuzz.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=300 \
  -k "sparse or csr or config4 or fuzz" 2>&1 | tail -5
timeout 300 python bench.py --workload config4 --steps 2 --warmup 1 > gpurun_out/bench_config4_$TAG.json 2> gpurun_out/bench_config4_$TAG.err
python - $TAG <<'PY'
import json,sys
d=json.loads([l for l in open("gpurun_out/bench_config4_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print(d["metric"], d["value"], d["ms_per_step"], d.get("breakdown_ms_per_step"))
PY
