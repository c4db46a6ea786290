#!/bin/bash
# warp-per-trajectory kernel with the software-pipelined row add: parity, bench sub-record; U = 16 variant
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=300 -k "generic or config1 or config2 or n24 or tiny_and_block or beyond or cli_gpu" 2>&1 | tail -3
for v in base; do
  LIB=""; [ $v != base ] && LIB=build/ab/$v/libonesolver_b200.so
  echo "== $v"
  [ $v != base ] && OSA_LIB_PATH=$LIB timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout=150 -k "generic_kernel or config1" 2>&1 | tail -1
  OSA_LIB_PATH=$LIB timeout 300 python bench.py --workload random_site --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(d['metric'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['achieved'], d['e2e']['value'])"
done
timeout 200 python tools/random_mode_probe.py 2>&1 | cut -c1-260
