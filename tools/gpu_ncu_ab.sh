#!/bin/bash
# ncu full captures of the dense sweep kernel, lock-step (OSA_WS_FLOW=0) and free-running (=1), on one
# wave of trajectories: tools/gpu_ncu_ab.sh TAG [SWEEPS]
TAG=${1:-x}; SW=${2:-2}
mkdir -p gpurun_out
for f in 0 1; do
  OSA_WS_FLOW=$f timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dense_seq -c 1 \
    -o gpurun_out/prof_flow${f}_$TAG -f python bench.py --steps 1 --warmup 0 --tries-per-gpu 1776 --sweeps $SW \
    --no-cpu-baseline --no-e2e > gpurun_out/ncu_flow${f}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_flow${f}_$TAG.log | cut -c1-200
done
