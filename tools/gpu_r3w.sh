#!/bin/bash
# after the allocation-order change: phases of the random-site end-to-end step, smoke, the dense parity tests, the default bench line
mkdir -p gpurun_out
python tools/e2e_random_probe.py 2>&1 | tail -10
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_pt.py tests/test_gpu_pa.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=300 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_final13.json 2> gpurun_out/bench_final13.err; tail -2 gpurun_out/bench_final13.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_final13.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["pageable_input"]["value"])
for o in d["other_configs"]:
    print(o["metric"], o["value"], o["ms_per_step"], o["roofline"]["frac"], o["e2e"]["value"])
PY
