#!/bin/bash
# end-of-round check on one GPU: smoke(), the whole GPU suite, the default bench line with its ncu launch list
TAG=${1:-final}
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider --tb=short --timeout=900 2>&1 | tail -8 > gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/ncu_launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e \
  > gpurun_out/ncu_launches_$TAG.log 2>&1
grep -c "k_" gpurun_out/ncu_launches_$TAG.csv
python - $TAG <<'PY'
import json, sys
d=json.loads([l for l in open("gpurun_out/bench_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
for o in d["other_configs"]:
    print(o["metric"], o["value"], o["ms_per_step"], o["roofline"]["frac"], o["e2e"]["value"])
PY
