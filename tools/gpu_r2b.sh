#!/bin/bash
# round-2 call B: the new GPU tests, racecheck of the N=4096 instantiations (one CTA), ncu evidence
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=600 \
  -k "csr_route or bench_schedule or config2" 2>&1 | tail -25 > gpurun_out/pytest_new_$TAG.log
tail -5 gpurun_out/pytest_new_$TAG.log
# racecheck, one CTA of the bench instantiation: free-running kernel (default) and lock-step kernel
for FLOW in 1 0; do
  OSA_WS_FLOW=$FLOW timeout 540 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 \
    python tools/race_n4096.py > gpurun_out/racecheck_n4096_flow${FLOW}_$TAG.log 2>&1
  echo "racecheck flow=$FLOW rc=$?"; tail -4 gpurun_out/racecheck_n4096_flow${FLOW}_$TAG.log | cut -c1-200
done
# ncu: launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/ncu_launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e \
  > gpurun_out/ncu_launches_$TAG.log 2>&1
tail -30 gpurun_out/ncu_launches_$TAG.csv | cut -c1-160
# full captures: sweep kernel (one wave, 2 sweeps), sparse kernel (config 4 shape, 4 sweeps), random-site kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dense_seq -c 1 \
  -o gpurun_out/prof_dense_seq_$TAG python bench.py --steps 1 --warmup 0 --tries-per-gpu 1776 --sweeps 2 \
  --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/ncu_full_seq_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_seq_$TAG.log | cut -c1-200
SWEEPS=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sparse -c 1 \
  -o gpurun_out/prof_sparse_$TAG python tools/sparse_once.py > gpurun_out/ncu_full_sparse_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_sparse_$TAG.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dense_generic -c 1 \
  -o gpurun_out/prof_generic_$TAG python tools/random_once.py > gpurun_out/ncu_full_generic_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_generic_$TAG.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
