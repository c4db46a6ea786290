#!/bin/bash
# random-site kernel, N=4096 fp32, whole waves for 12 and 13 warps per SM (23088 = 148 x 12 x 13 tries):
# flip-by-flip kernel / batched rows with one buffer (6b46993) / two buffers + unpredicated pieces, CTA sizes
TAG=${1:-r3s}
mkdir -p gpurun_out
run() { # label libpath wpb
  OSA_LIB_PATH=$2 OSA_GEN_WPB=$3 TRIES=${TRIES:-23088} ITERS=${ITERS:-2048} timeout 300 python tools/random_mode_probe.py 2>&1 | grep "^{" | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('== $1 wpb=${3:-auto}', '%.4e attempts/s' % d['attempts_per_s'], '%.0f GB/s' % d['row_gbs'], 'acc %.6f' % d['accept_frac'], d['ms_sweep'], 'ms')"
}
run flip_by_flip build/ab/gen_old/libonesolver_b200.so ""
run one_buffer build/ab/g_head/libonesolver_b200.so ""
for w in "" 4 3 2 1; do run tree "" "$w"; done
for d in $(cd build/ab && ls -d g_x* 2>/dev/null); do run $d build/ab/$d/libonesolver_b200.so ""; run $d build/ab/$d/libonesolver_b200.so 4; done
