"""BASELINE.json configs 2, 3 and 4 at their full sizes on one GPU: wall numbers for DESIGN.md.
(config 1 is the CPU CLI run of tests/test_host_cpp.py, config 5 is bench.py.)"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from onesolver_b200 import Problem, capi, exhaustive  # noqa: E402
from onesolver_b200 import problems as gen  # noqa: E402


def geo(n, lo, hi):
    return lo * (hi / lo) ** (np.arange(n) / max(1, n - 1))


def report(name, res, extra=None):
    st = res.stats
    out = {"config": name, "attempts": st["attempts"], "ms_sweep": round(st["ms_sweep"], 2),
           "ms_energy": round(st["ms_energy"], 2), "ms_total": round(st["ms_total"], 2),
           "attempts_per_s": st["attempts"] / (st["ms_sweep"] * 1e-3),
           "accept_frac": round(st["accepts"] / st["attempts"], 4), "kernel": st["kernel_id"],
           "traj_per_row_fetch": st["traj_per_batch"], "best_energy": res.energy}
    out.update(extra or {})
    print(json.dumps(out), flush=True)


# config 2: dense N=24, 4096 tries, reference mode, best state vs exhaustive ground state
q = gen.dense_integer_qubo(24, seed=2026)
with Problem.dense(q) as p:
    for _ in range(2):
        r = p.anneal(geo(400, 0.5, 20.0), 400, 4096, mode=capi.MODE_RANDOM_SITE)
state, e0 = exhaustive(q)
report("2: dense N=24, 4096 tries, 400 random-site iterations (reference mode)", r,
       {"exhaustive_ground": e0, "ground_state_found": bool(r.energy == e0)})

# config 3: dense fp64 N=1024, 16384 tries, 1000 sweeps
q = gen.dense_uniform_qubo(1024, seed=2027)
s = np.sqrt(1024)
with Problem.dense(q, sweep_precision=capi.SWEEP_F64) as p:
    t0 = time.perf_counter()
    r = p.anneal(geo(1000, 0.3 * s, 0.02 * s), 1000, 16384, mode=capi.MODE_SEQUENTIAL_SWEEP)
    wall = time.perf_counter() - t0
bytes_rows = (r.stats["row_fetches"] + r.stats["init_row_fetches"]) * 1024 * 8
report("3: dense fp64 N=1024, 16384 tries, 1000 sweeps", r,
       {"wall_s": round(wall, 3), "row_gbs": bytes_rows / (r.stats["ms_sweep"] * 1e-3) / 1e9})

# config 4: sparse Pegasus-like N=5627 (degree <= 15), 65536 tries, linear schedule, 100 sweeps
n = 5627
rowptr, col, val, diag = gen.sparse_random_graph(n, 15, seed=2028)
for prec, name in ((capi.SWEEP_F32, "f32"), (capi.SWEEP_F64, "f64")):
    with Problem.csr(rowptr, col, val, diag, sweep_precision=prec) as p:
        t0 = time.perf_counter()
        r = p.anneal(np.linspace(0.5, 5.0, 100), 100, 65536, mode=capi.MODE_SEQUENTIAL_SWEEP)
        wall = time.perf_counter() - t0
    report("4: sparse N=5627 CSR (%d couplers), 65536 tries, linear schedule, 100 sweeps, %s" %
           (len(col) // 2, name), r, {"wall_s": round(wall, 3)})
