#!/bin/bash
# round-2 call E: sparse kernel v3 (groups of 4/8, batched decisions): parity, bench line of config 4
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=600 \
  -k "sparse or csr or config4" 2>&1 | tail -25 > gpurun_out/pytest_sparse_$TAG.log
tail -5 gpurun_out/pytest_sparse_$TAG.log
timeout 600 python bench.py --workload config4 --steps 2 --warmup 1 > gpurun_out/bench_config4_$TAG.json 2> gpurun_out/bench_config4_$TAG.err
python - <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1] if len(sys.argv)>1 else "gpurun_out/bench_config4_'"$TAG"'.json") if l.startswith("{")][-1])
print(d["metric"], d["value"], d["ms_per_step"], d.get("breakdown_ms_per_step"))
PY
SWEEPS=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sparse -c 1 \
  -o gpurun_out/prof_sparse_$TAG python tools/sparse_once.py > gpurun_out/ncu_full_sparse_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_sparse_$TAG.log | cut -c1-200
