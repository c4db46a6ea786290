#!/bin/bash
# round-2 call A: parity (all GPU tests), TMEM probe, bench line
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider --tb=short --timeout=600 2>&1 | tail -25 > gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 120 ./build/bin/microbench_tmem 4000 > gpurun_out/microbench_tmem_$TAG.log 2>&1; cat gpurun_out/microbench_tmem_$TAG.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-600 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
