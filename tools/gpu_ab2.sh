#!/bin/bash
# same-box A/B of the sweep kernel without the parity run: tools/probe.py dense for LABEL=PATH
# libraries and the in-tree build, interleaved twice (run-to-run noise), then the cp.async probes
TAG=${1:-ab}; shift
mkdir -p gpurun_out
for rep in 1 2; do
  for kv in "$@"; do
    label=${kv%%=*}; path=${kv#*=}
    OSA_LIB_PATH=$path timeout 300 python tools/probe.py dense > gpurun_out/probe_${TAG}_${label}_$rep.log 2>&1
  done
  timeout 300 python tools/probe.py dense > gpurun_out/probe_${TAG}_tree_$rep.log 2>&1
done
timeout 120 ./build/bin/microbench 16384 4096 2 | grep cpasync > gpurun_out/microbench_$TAG.log 2>&1
for f in gpurun_out/probe_${TAG}_*.log; do echo "== $f"; python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['probe'], d['ms_sweep'], '%.3e' % d['attempts_per_s'], d['kcyc_per_cta'])
PY
done
cat gpurun_out/microbench_$TAG.log
