#!/bin/bash
# bench (both arms) + ncu launch list + one full ncu capture of the sweep kernel + CLI on the GPU
mkdir -p gpurun_out
TAG=${1:-v1}
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
# launch list of the same command (short): per-launch times are cold-cache/serialised -> shares only
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --tries-per-gpu 9472 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_$TAG.log 2>&1
# full capture of the sweep kernel on a one-wave problem
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dense_seq -c 1 \
  -o gpurun_out/prof_dense_seq_$TAG python bench.py --steps 1 --warmup 0 --tries-per-gpu 1776 --sweeps 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
./build/bin/one-solver-anneal --input examples/test1.qubo --output gpurun_out/cli_gpu_test1.csv --device-type gpu --stats > gpurun_out/cli_gpu.log 2>&1
cat gpurun_out/cli_gpu_test1.csv >> gpurun_out/cli_gpu.log
./build/bin/one-solver-anneal --input tests/golden/chimera512/001.qubo --output gpurun_out/cli_gpu_c512.csv --device-type gpu --num-iter 1000 --num-tries 460 --beta-max 10 --stats >> gpurun_out/cli_gpu.log 2>&1
tail -c 300 gpurun_out/cli_gpu_c512.csv >> gpurun_out/cli_gpu.log
cat gpurun_out/bench_$TAG.json; cat gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_$TAG.err; tail -12 gpurun_out/cli_gpu.log
