#!/bin/bash
# same-box A/B, one pass: tools/probe.py dense for LABEL=PATH libraries and the in-tree build
TAG=${1:-ab}; shift
mkdir -p gpurun_out
for kv in "$@" tree=; do
  label=${kv%%=*}; path=${kv#*=}
  OSA_LIB_PATH=$path timeout 300 python tools/probe.py dense > gpurun_out/probe_${TAG}_$label.log 2>&1
  echo "== $label"
  python - gpurun_out/probe_${TAG}_$label.log <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["probe"], d["ms_sweep"], "%.3e" % d["attempts_per_s"], d["kcyc_per_cta"])
PY
done
