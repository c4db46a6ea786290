#!/bin/bash
# does the poll interval change the number of executed instructions (and the time) of the free-running kernel?
mkdir -p gpurun_out
for v in base sleep400 sleep2000; do
  LIB=""; [ $v != base ] && LIB=build/ab/$v/libonesolver_b200.so
  echo "== $v"
  OSA_LIB_PATH=$LIB timeout 200 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed \
     --clock-control none -k regex:k_dense_seq -c 1 python tools/flow_once.py 2>&1 | grep -E "smsp__inst|gpu__time|per_cycle|rows"
done 2>&1 | tee gpurun_out/poll_interval_instruction_counts.txt
