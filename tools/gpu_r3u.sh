#!/bin/bash
# threshold of the batched row add: flip-by-flip kernel (gen_old) against the in-tree library over row lengths
mkdir -p gpurun_out
export CASES="2048:7104:4096:f64,4096:3552:4096:f64,2048:7104:8192:f32,3072:7104:4096:f32,8192:3552:2048:f32,6000:3552:4096:f32"
for kv in old=build/ab/gen_old/libonesolver_b200.so new=; do
  label=${kv%%=*}; path=${kv#*=}
  echo "== $label"
  OSA_LIB_PATH=$path timeout 300 python tools/random_mode_probe.py 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   n=%d %s %.4e attempts/s %.0f GB/s acc %.6f' % (d['n'], 'tries=%d' % d['tries'], d['attempts_per_s'], d['row_gbs'], d['accept_frac']))"
done
