#!/bin/bash
# Build a variant of the sweep kernels next to the in-tree library, for same-box A/B runs
# (tools/gpu_ab3.sh / gpu_ab4.sh label=build/ab/NAME/libonesolver_b200.so):
#   tools/build_variant.sh NAME [-DMACRO[=VALUE] ...]
# Variants are probe builds (-DOSA_PROBE): they honour the timing-experiment switches OSA_WS_DEBUG /
# OSA_DS_CFG=81211 that the in-tree library ignores (osa_common.cuh, probe_env_int).
# compiles osa_dense_seq_ws.cu, osa_dense_seq_ws2.cu and osa_dense_seq.cu from the working tree with the given macros into
# build/ab/NAME/ and links them with the other objects of the last in-tree build (build/csrc).
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/build/ab/$NAME; O=$ROOT/build/csrc
mkdir -p "$OUT"
NVFLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr -ccbin /usr/bin/g++ -DOSA_PROBE"
for src in osa_dense_seq_ws osa_dense_seq_ws2 osa_dense_seq; do
  (cd "$ROOT/onesolver_b200/csrc" && nvcc $NVFLAGS "$@" -c $src.cu -o "$OUT/$src.o" 2> "$OUT/$src.ptxas.log") &
done
wait
grep -il " error" "$OUT"/*.ptxas.log && exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libonesolver_b200.so" "$O/osa_api.o" \
  "$OUT/osa_dense_seq.o" "$OUT/osa_dense_seq_ws.o" "$OUT/osa_dense_seq_ws2.o" "$O/osa_dense_generic.o" "$O/osa_dense_init.o" "$O/osa_sparse.o" \
  "$O/osa_energy.o" "$O/osa_exhaustive.o" "$O/osa_pt.o" "$O/osa_pa.o" "$O/osa_multi.o" -ldl -ccbin /usr/bin/g++
echo "spills per instantiation (count, ptxas line):"
grep -h "bytes spill" "$OUT/osa_dense_seq_ws.ptxas.log" "$OUT/osa_dense_seq_ws2.ptxas.log" | sort | uniq -c
ls -la "$OUT/libonesolver_b200.so"
