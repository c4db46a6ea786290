"""Per-source-line totals (instructions executed, stall samples) from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`: where the instructions of a kernel go."""
import collections
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
inst, samp, text = collections.Counter(), collections.Counter(), {}
cur_file, cur = "?", None
hdr = None
for r in csv.reader(open(path)):
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        cur = (cur_file, int(r[0]))
        text[cur] = r[1].strip()
        continue
    if cur is None:
        continue
    try:
        inst[cur] += int(r[ie] or 0)
        samp[cur] += int(r[isamp] or 0)
    except ValueError:
        pass
ti, ts = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {ti}, samples {ts}")
for k, v in inst.most_common(top):
    print(f"{100.0 * v / ti:5.1f}% inst {100.0 * samp[k] / max(ts, 1):5.1f}% samp  {k[0]}:{k[1]:<4d} {text[k][:100]}")
