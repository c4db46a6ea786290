#!/bin/bash
# SASS of the N = 4096 fp32 instantiation of k_dense_seq_ws alone (plain annealing, not the PT one):
#   tools/sass_ws4096.sh OUT.sass [-DMACRO ...]
set -e
OUT=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
(cd "$ROOT/onesolver_b200/csrc" && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
  -Xptxas -v --expt-relaxed-constexpr -ccbin /usr/bin/g++ -DOSA_WS_ONLY_F32_4 "$@" -c osa_dense_seq_ws.cu \
  -o "$TMP/ws.o" 2> "$TMP/ptxas.log")
grep -A1 "Lb0ELi4EE" "$TMP/ptxas.log" | grep "registers\|spill" || true
cuobjdump -sass "$TMP/ws.o" | awk '/Function :/{f = ($0 ~ /Lb0ELi4EE/)} f' | grep -v "^\s*/\* 0x" | sed 's/\/\* 0x[0-9a-f]* \*\///' > "$OUT"
wc -l "$OUT"
rm -rf "$TMP"
