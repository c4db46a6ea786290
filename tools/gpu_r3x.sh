#!/bin/bash
# compute-sanitizer on the warp-per-trajectory kernel (flip-by-flip and batched row add, both precisions, ragged rows)
mkdir -p gpurun_out
SEL='test_dense_generic_row_add_forms_bit_exact or (test_dense_generic_kernel_bit_exact and (150 or 1100))'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --tb=line --timeout=800 \
    -k "$SEL" > gpurun_out/sanitizer_generic_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_generic_$tool.log | tail -6
done
