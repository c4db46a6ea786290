#!/bin/bash
# same-box A/B over environment knobs of the in-tree library: for every VAR=VALUE argument (use
# NONE=0 for the default) run tools/probe.py dense and a short bench-shaped run (4 waves, 32 sweeps)
TAG=${1:-env}; shift
mkdir -p gpurun_out
for kv in "$@"; do
  label=${kv//=/}
  env $kv timeout 300 python tools/probe.py dense > gpurun_out/probe_${TAG}_$label.log 2>&1
  env $kv timeout 300 python bench.py --steps 2 --warmup 1 --tries-per-gpu 7104 --no-cpu-baseline --no-e2e \
    > gpurun_out/bench_${TAG}_$label.json 2> gpurun_out/bench_${TAG}_$label.err
  echo "== $kv"
  python - gpurun_out/probe_${TAG}_$label.log gpurun_out/bench_${TAG}_$label.json <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['probe'], d['ms_sweep'], '%.3e' % d['attempts_per_s'], d['kcyc_per_cta'])
for l in open(sys.argv[2]):
    if l.startswith('{'):
        d = json.loads(l)
        print('bench-shaped', '%.4e' % d['value'], d['breakdown_ms_per_step'])
PY
done
