#!/bin/bash
# same-box A/B over probe builds (tools/build_variant.sh): tools/gpu_ab5.sh TAG "MODES" LABEL...
# runs `tools/probe.py MODES` against build/ab/LABEL/libonesolver_b200.so for every label and prints a
# one-line summary per probe; logs go to gpurun_out/probe_TAG_LABEL.log
TAG=$1; MODES=$2; shift 2
mkdir -p gpurun_out
for label in "$@"; do
  OSA_LIB_PATH=build/ab/$label/libonesolver_b200.so timeout 600 python tools/probe.py $MODES \
    > gpurun_out/probe_${TAG}_$label.log 2>&1
  echo "== $label"
  python - gpurun_out/probe_${TAG}_$label.log <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        extra = " frac=%.4f" % d["frac_of_18223"] if "frac_of_18223" in d else ""
        print(d["probe"], d["ms_sweep"], "%.3e" % d["attempts_per_s"], "acc=%.3f" % d.get("accept_frac", 0),
              d["kcyc_per_cta"], "rows=%d" % d.get("row_fetches", 0), extra)
    elif "rror" in l:
        print(l.rstrip())
PY
done
