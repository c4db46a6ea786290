#!/bin/bash
# SASS of the N = 4096 fp32 instantiation of k_dense_seq_flow (plain annealing) and a summary of its
# first row loop: tools/sass_flow4096.sh OUT.sass [-DMACRO ...]
set -e
OUT=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
(cd "$ROOT/onesolver_b200/csrc" && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
  -Xptxas -v --expt-relaxed-constexpr -ccbin /usr/bin/g++ -DOSA_WS_ONLY_F32_4 "$@" -c osa_dense_seq_ws2.cu \
  -o "$TMP/ws.o" 2> "$TMP/ptxas.log")
grep -A1 "Lb0EEE" "$TMP/ptxas.log" | grep "spill" || true
cuobjdump -sass "$TMP/ws.o" | awk '/Function :/{f = ($0 ~ /Lb0EEE/)} f' | grep -v "^\s*/\* 0x" | sed 's/\/\* 0x[0-9a-f]* \*\///' > "$OUT"
python3 - "$OUT" <<'PY'
import sys, re
lines = open(sys.argv[1]).read().split("\n")
idx = [i for i, l in enumerate(lines) if "DEPBAR.LE SB0, 0xa" in l]
for k, i in enumerate(idx[:2]):
    # loop = from the nearest preceding branch target... approximate: up to the next backward branch
    j = i
    while j < len(lines) and not re.search(r"BRA(\.U)? .*0x", lines[j]) or (j < len(lines) and int(re.search(r"0x([0-9a-f]+) ;", lines[j]).group(1), 16) > int(re.search(r"/\*([0-9a-f]+)\*/", lines[j]).group(1), 16)):
        j += 1
    body = lines[i - 6:j + 1]
    ops = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s*(@!?U?P\d\s+)?", "", l).split(" ")[0] for l in body]
    from collections import Counter
    c = Counter(o.split(".")[0] for o in ops)
    print("loop %d: %d instructions; " % (k, len(body)), dict(c.most_common(12)))
PY
rm -rf "$TMP"
