#!/bin/bash
# GPU check of the free-running dense kernel: memcheck of smoke(), parity tests (in-tree library,
# OSA_WS_FLOW=1 selects the kernel), then the probes of the probe builds given as arguments
# (build/ab/LABEL, each with staged rows and with the ring: OSA_FLOW_STAGE=1/0)
mkdir -p gpurun_out
echo "== memcheck smoke, free-running kernel"
OSA_WS_FLOW=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
echo "== pytest, free-running kernel, staged rows"; OSA_WS_FLOW=1 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
echo "== pytest, free-running kernel, ring"; OSA_WS_FLOW=1 OSA_FLOW_STAGE=0 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for label in "$@"; do
for st in 1 0; do
  echo "== probes $label OSA_FLOW_STAGE=$st"
  OSA_LIB_PATH=build/ab/$label/libonesolver_b200.so OSA_FLOW_STAGE=$st timeout 600 python tools/probe.py dense benchlike > gpurun_out/probe_${label}_st$st.log 2>&1
  python - gpurun_out/probe_${label}_st$st.log <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        extra = " frac=%.4f" % d["frac_of_18223"] if "frac_of_18223" in d else ""
        print(d["probe"], "R=%s" % d.get("R"), d["ms_sweep"], "%.3e" % d["attempts_per_s"], "acc=%.3f" % d.get("accept_frac", 0),
              d["kcyc_per_cta"], "rows=%d" % d.get("row_fetches", 0), extra)
    elif "rror" in l:
        print(l.rstrip())
PY
done
done
