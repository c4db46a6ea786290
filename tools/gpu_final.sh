#!/bin/bash
# round-end validation on one GPU: parity tests, both bench arms, configs 3 and 4 through bench.py
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --tb=short --timeout=240 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
timeout 600 python bench.py --workload config3 --steps 2 --warmup 3 > gpurun_out/bench_config3_$TAG.json 2>> gpurun_out/bench_$TAG.err
timeout 600 python bench.py --workload config4 --steps 2 --warmup 3 > gpurun_out/bench_config4_$TAG.json 2>> gpurun_out/bench_$TAG.err
timeout 600 python bench.py --workload config4 --precision f64 --steps 2 --warmup 3 > gpurun_out/bench_config4_f64_$TAG.json 2>> gpurun_out/bench_$TAG.err
for f in bench bench_ref bench_config3 bench_config4 bench_config4_f64; do echo "== $f"; cut -c1-1800 gpurun_out/${f}_$TAG.json; done
tail -5 gpurun_out/bench_$TAG.err
