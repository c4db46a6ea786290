#!/bin/bash
# random-site kernel: grouped refills (UG), loads in flight (U), walk of batch b+1 over the stream of
# batch b (AHEAD): same-box A/B of the random-mode probe; parity of one AHEAD variant
TAG=${1:-r3o}
mkdir -p gpurun_out
for d in gen_old "" $(cd build/ab && ls -d g_*); do
  label=${d:-tree}; path=${d:+build/ab/$d/libonesolver_b200.so}
  OSA_LIB_PATH=$path timeout 300 python tools/random_mode_probe.py > gpurun_out/random_probe_${TAG}_$label.log 2>&1
  echo "== $label $(python - gpurun_out/random_probe_${TAG}_$label.log <<'PY'
import sys, json
out = []
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); out.append("n=%d %.3e (%.0f GB/s, acc %.6f)" % (d["n"], d["attempts_per_s"], d["row_gbs"], d["accept_frac"]))
print(" | ".join(out) if out else open(sys.argv[1]).read()[-300:])
PY
)"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=300 \
  -k "not sparse and not csr and not cli" 2>&1 | tail -4
