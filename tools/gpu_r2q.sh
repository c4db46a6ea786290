#!/bin/bash
# free-running kernel with mbarrier hand-over (build/ab/mbar1) against the in-tree polling version: parity, time, instruction count
mkdir -p gpurun_out
OSA_LIB_PATH=build/ab/mbar1/libonesolver_b200.so timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=60 -k "single_role or dense_seq_bit_exact" 2>&1 | tail -4
for v in base mbar1; do
  LIB=""; [ $v != base ] && LIB=build/ab/$v/libonesolver_b200.so
  echo "== $v"
  OSA_LIB_PATH=$LIB timeout 100 python tools/flow_once.py | cut -c1-200
  OSA_LIB_PATH=$LIB timeout 100 python tools/flow_once.py | cut -c1-200
  OSA_LIB_PATH=$LIB timeout 200 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed \
     --clock-control none -k regex:k_dense_seq -c 1 python tools/flow_once.py 2>&1 | grep -E "smsp__inst|gpu__time|per_cycle"
done 2>&1 | tee gpurun_out/ab_mbarrier_handover.txt
bash tools/gpu_ab5.sh r2q "dense4k benchlike" mbar1 2>&1 | tee -a gpurun_out/ab_mbarrier_handover.txt
