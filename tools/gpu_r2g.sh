#!/bin/bash
# round-2 call G (2 GPUs): multi-device tests, 2-GPU bench line, ncu capture of the config-3 kernel
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_multi.py -q -x -p no:cacheprovider --tb=short --timeout=600 2>&1 | tail -15 > gpurun_out/pytest_multi_$TAG.log
tail -4 gpurun_out/pytest_multi_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu_$TAG.json 2> gpurun_out/bench_2gpu_$TAG.err
grep '^{' gpurun_out/bench_2gpu_$TAG.json | cut -c1-200; tail -2 gpurun_out/bench_2gpu_$TAG.err
CUDA_VISIBLE_DEVICES=0 SWEEPS=20 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dense_seq -c 1 \
  -o gpurun_out/prof_config3_$TAG python tools/config3_once.py > gpurun_out/ncu_full_config3_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_config3_$TAG.log | cut -c1-200
