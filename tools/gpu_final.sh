#!/bin/bash
# end-of-round check on one GPU: smoke(), the whole GPU suite, the default bench line and the reference arm
TAG=${1:-final}
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider --tb=short --timeout=900 2>&1 | tail -8 > gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python - $TAG <<'PY'
import json, sys
d=json.loads([l for l in open("gpurun_out/bench_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
for o in d["other_configs"]:
    print(o["metric"], o["value"], o["ms_per_step"], o["roofline"]["frac"], o["e2e"]["value"])
r=json.loads([l for l in open("gpurun_out/bench_ref_%s.json" % sys.argv[1]) if l.startswith("{")][-1])
print("reference arm:", r["value"], r["cpu_baseline"]["cores"], r["ms_per_step"])
PY
