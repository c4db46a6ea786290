#!/bin/bash
# round-2 call D: poll interval of the free-running kernel (A/B), ncu capture of the grouped sparse kernel
TAG=${1:-r2d}
mkdir -p gpurun_out
bash tools/gpu_ab5.sh $TAG "dense4k benchlike" sleep100 sleep400 sleep1500 > gpurun_out/ab_$TAG.txt 2>&1; cat gpurun_out/ab_$TAG.txt
SWEEPS=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sparse -c 1 \
  -o gpurun_out/prof_sparse_$TAG python tools/sparse_once.py > gpurun_out/ncu_full_sparse_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_sparse_$TAG.log | cut -c1-200
