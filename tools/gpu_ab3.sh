#!/bin/bash
# parity tests with the in-tree library, then tools/probe.py dense for LABEL=PATH libraries and
# the in-tree build; one summary line per probe
TAG=${1:-ab}
bash tools/gpu_ab.sh "$@" > /dev/null 2>&1
tail -4 gpurun_out/pytest_$TAG.log
for f in gpurun_out/probe_${TAG}_*.log; do
  echo "== $f"
  python - "$f" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["probe"], d["ms_sweep"], "%.3e" % d["attempts_per_s"], d["kcyc_per_cta"])
PY
done
