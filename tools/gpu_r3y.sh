#!/bin/bash
# CTA size of the warp-per-trajectory kernel by occupancy: four-warp CTAs (OSA_GEN_WPB=4, the old rule) against the pick
mkdir -p gpurun_out
export CASES="4096:3552:4096:f64,8192:3552:2048:f32,6000:5328:4096:f32,2048:7104:4096:f64,4096:7104:4096:f32,1024:14208:8192:f64,128:65536:4096:f64"
for w in 4 ""; do
  echo "== wpb=${w:-occupancy}"
  OSA_GEN_WPB=$w timeout 300 python tools/random_mode_probe.py 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   n=%d %s %.4e attempts/s %.0f GB/s acc %.6f' % (d['n'], 'tries=%d' % d['tries'], d['attempts_per_s'], d['row_gbs'], d['accept_frac']))"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --tb=short --timeout=300 -k "generic or config1 or config2" 2>&1 | tail -2
