#!/bin/bash
# one optimisation iteration on the GPU: parity first, then numbers, optionally an ncu capture
TAG=${1:-iter}; PROF=${2:-0}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --tb=short --timeout=240 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python tools/probe.py dense sparse > gpurun_out/probe_$TAG.log 2>&1; cat gpurun_out/probe_$TAG.log
if [ "$PROF" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dense_seq -c 1 \
    -o gpurun_out/prof_dense_seq_$TAG python bench.py --steps 1 --warmup 0 --tries-per-gpu 1776 --sweeps 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-300
fi
if [ "$PROF" = "2" ]; then
  timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json | cut -c1-2500
fi
