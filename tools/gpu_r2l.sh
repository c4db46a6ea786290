#!/bin/bash
# round-2 call L: config-3 shape (dense fp64 N=1024) with two CTAs per SM (OSA_WS_C3=1/2) against the default
TAG=${1:-r2l}
mkdir -p gpurun_out
for c in 0 1 2; do
  echo "== OSA_WS_C3=$c"
  OSA_WS_C3=$c SWEEPS=100 timeout 60 python tools/config3_once.py || echo "FAILED/timeout rc=$?"
  OSA_WS_C3=$c SWEEPS=100 timeout 60 python tools/config3_once.py || echo "FAILED/timeout rc=$?"
done 2>&1 | tee gpurun_out/ab_config3_two_ctas_$TAG.txt
for c in 1 2; do
  OSA_WS_C3=$c timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=60 -k "dense_seq_bit_exact or sweeps_per_beta" 2>&1 | tail -3
done
