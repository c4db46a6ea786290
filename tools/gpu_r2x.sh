#!/bin/bash
# lock-step kernel with chunked tile loads in the catch-up of the decide role (build/ab/chunk2, chunk4) vs in-tree
mkdir -p gpurun_out
for v in base chunk2 chunk4; do
  LIB=""; [ $v != base ] && LIB=build/ab/$v/libonesolver_b200.so
  echo "== $v"
  [ $v != base ] && OSA_LIB_PATH=$LIB timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout=100 -k "dense_seq_bit_exact or single_role" 2>&1 | tail -1
  OSA_LIB_PATH=$LIB SWEEPS=100 timeout 60 python tools/config3_once.py
  OSA_LIB_PATH=$LIB SWEEPS=100 timeout 60 python tools/config3_once.py
  OSA_LIB_PATH=$LIB OSA_WS_FLOW=0 timeout 100 python tools/flow_once.py | cut -c1-200
done 2>&1 | tee gpurun_out/ab_catchup_chunk.txt
