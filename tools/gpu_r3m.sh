#!/bin/bash
# random-site kernel with batched row adds: parity (dense tests), then same-box A/B of the
# random-mode probe: old kernel / new (U=8, in-tree) / U=12 / U=16
TAG=${1:-r3m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=300 \
  -k "not sparse and not csr and not cli" 2>&1 | tail -8
for kv in old=build/ab/gen_old/libonesolver_b200.so new= u12=build/ab/gen_u12/libonesolver_b200.so u16=build/ab/gen_u16/libonesolver_b200.so; do
  label=${kv%%=*}; path=${kv#*=}
  echo "== $label"
  OSA_LIB_PATH=$path timeout 300 python tools/random_mode_probe.py 2>&1 | tee gpurun_out/random_probe_${TAG}_$label.log | cut -c1-230
done
