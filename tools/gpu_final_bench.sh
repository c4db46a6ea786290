#!/bin/bash
# the default bench line and its ncu launch list (after tools/gpu_final.sh has run the suite)
TAG=${1:-final}
mkdir -p gpurun_out
( time timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/ncu_launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e \
  > gpurun_out/ncu_launches_$TAG.log 2>&1
grep -c "k_" gpurun_out/ncu_launches_$TAG.csv
( time timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err ) 2>&1 | grep real; cut -c1-400 gpurun_out/bench_ref_$TAG.json
python - $TAG <<'PY'
import json, sys
d=json.loads([l for l in open("gpurun_out/bench_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
for o in d["other_configs"]:
    print(o["metric"], o["value"], o["ms_per_step"], o["roofline"]["frac"], o["e2e"]["value"], o["steps"], o["warmup"])
PY
