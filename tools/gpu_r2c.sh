#!/bin/bash
# round-2 call C: all GPU tests (population annealing, grouped sparse layout, variant test), racecheck
# of the lock-step kernel after the idle-lane fix, bench line
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider --tb=short --timeout=900 2>&1 | tail -40 > gpurun_out/pytest_$TAG.log
tail -12 gpurun_out/pytest_$TAG.log
OSA_WS_FLOW=0 timeout 540 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 \
  python tools/race_n4096.py > gpurun_out/racecheck_n4096_flow0_$TAG.log 2>&1
echo "racecheck lock-step rc=$?"; tail -3 gpurun_out/racecheck_n4096_flow0_$TAG.log | cut -c1-200
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-300 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_r2c.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"])
for o in d["other_configs"]:
    print(o["metric"], o["value"], o["ms_per_step"], o["roofline"]["frac"])
PY
