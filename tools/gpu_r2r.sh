#!/bin/bash
# free-running kernel with one poller per role (build/ab/onepoll) against the in-tree version
mkdir -p gpurun_out
OSA_LIB_PATH=build/ab/onepoll/libonesolver_b200.so timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=60 -k "single_role or dense_seq_bit_exact" 2>&1 | tail -4
for v in base onepoll; do
  LIB=""; [ $v != base ] && LIB=build/ab/$v/libonesolver_b200.so
  echo "== $v"
  OSA_LIB_PATH=$LIB timeout 100 python tools/flow_once.py | cut -c1-200
  OSA_LIB_PATH=$LIB timeout 100 python tools/flow_once.py | cut -c1-200
  OSA_LIB_PATH=$LIB timeout 200 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed \
     --clock-control none -k regex:k_dense_seq -c 1 python tools/flow_once.py 2>&1 | grep -E "smsp__inst|gpu__time|per_cycle"
done 2>&1 | tee gpurun_out/ab_one_poller.txt
bash tools/gpu_ab5.sh r2r "dense4k benchlike" onepoll 2>&1 | tee -a gpurun_out/ab_one_poller.txt
