#!/bin/bash
# Variant of the warp-per-trajectory kernel (osa_dense_generic.cu) next to the in-tree library for
# same-box A/B runs: tools/build_variant_gen.sh NAME [-DOSA_GEN_U=.. -DOSA_GEN_UG=.. -DOSA_GEN_AHEAD=.. ...]
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/build/ab/$NAME; O=$ROOT/build/csrc
mkdir -p "$OUT"
(cd "$ROOT/onesolver_b200/csrc" && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v \
  --expt-relaxed-constexpr -ccbin /usr/bin/g++ "$@" -c osa_dense_generic.cu -o "$OUT/osa_dense_generic.o" 2> "$OUT/ptxas.log")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libonesolver_b200.so" "$O/osa_api.o" "$O/osa_dense_seq.o" \
  "$O/osa_dense_seq_ws.o" "$O/osa_dense_seq_ws2.o" "$OUT/osa_dense_generic.o" "$O/osa_dense_init.o" "$O/osa_sparse.o" \
  "$O/osa_energy.o" "$O/osa_exhaustive.o" "$O/osa_pt.o" "$O/osa_pa.o" "$O/osa_multi.o" -ldl -ccbin /usr/bin/g++
echo "$NAME: $(grep -A1 'Lb1ELb1E' "$OUT/ptxas.log" | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"
