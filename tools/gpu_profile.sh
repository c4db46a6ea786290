#!/bin/bash
# ncu evidence for one round: launch list of the bench command, full captures of the sweep kernel
# (one wave) and of the DMMA energy kernel, and L2/DRAM traffic of the sweep kernel at bench size.
TAG=${1:-r}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/ncu_launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e \
  > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -12 gpurun_out/ncu_launches_$TAG.csv | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dense_seq -c 1 \
  -o gpurun_out/prof_dense_seq_$TAG python bench.py --steps 1 --warmup 0 --tries-per-gpu 1776 --sweeps 2 \
  --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_seq_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_seq_$TAG.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_energy_dense_mma -c 1 \
  -o gpurun_out/prof_energy_mma_$TAG python bench.py --steps 1 --warmup 0 --tries-per-gpu 37888 --sweeps 2 \
  --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_energy_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_energy_$TAG.log | cut -c1-200
timeout 1200 ncu --metrics lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none -k regex:k_dense_seq -c 1 --csv --log-file gpurun_out/ncu_traffic_$TAG.csv \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_traffic_$TAG.log 2>&1
tail -8 gpurun_out/ncu_traffic_$TAG.csv | cut -c1-300
