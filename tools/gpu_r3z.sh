#!/bin/bash
# row-add form again, now with the CTA size by occupancy (6 warps per SM at 32 KiB fields): flip by flip against batched
mkdir -p gpurun_out
export CASES="4096:3552:4096:f64,8192:3552:2048:f32,6000:5328:4096:f32,2048:7696:4096:f64,5000:6512:4096:f32"
for b in 0 1; do
  echo "== OSA_GEN_BATCH=$b"
  OSA_GEN_BATCH=$b timeout 300 python tools/random_mode_probe.py 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   n=%d %s %.4e attempts/s %.0f GB/s acc %.6f' % (d['n'], 'tries=%d' % d['tries'], d['attempts_per_s'], d['row_gbs'], d['accept_frac']))"
done
