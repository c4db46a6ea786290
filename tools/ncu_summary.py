"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv`."""
import collections
import csv
import re
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda name: vals[hdr.index(name)] if name in hdr else "n/a"
for name in ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread",
             "launch__grid_size", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
             "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
             "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
             "dram__bytes_write.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
             "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
             "smsp__warps_active.avg.per_cycle_active", "l1tex__t_sector_hit_rate.pct",
             "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]:
    print(f"{name:70s} {get(name)}")
print("-- stall reasons (warps per issue-active cycle)")
for i, h in enumerate(hdr):
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        try:
            v = float(vals[i])
        except ValueError:
            continue
        if v > 0.03:
            print(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:.3f}")
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = {k: hdr.index(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
ops, samp, data = collections.Counter(), collections.Counter(), []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
    op = m.group(2).split(".")[0] if m else "?"
    n, s = int(r[ie] or 0), int(r[isamp] or 0)
    ops[op] += n
    samp[op] += s
    data.append((s, n, r[ia], {k: int(r[v] or 0) for k, v in stall_cols.items()}))
tot, ts = sum(ops.values()), sum(samp.values())
print(f"-- instruction mix: {tot} warp instructions, {ts} samples")
for op, n in ops.most_common(14):
    print(f"  {op:10s} inst {n / tot * 100:5.1f}%  samples {samp[op] / ts * 100:5.1f}%")
print("-- hottest instructions")
for s, n, text, d in sorted(data, key=lambda x: -x[0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]:
    top = sorted(d.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {s / ts * 100:5.2f}% n={n:9d} {text.strip()[:64]:64s} {top}")
