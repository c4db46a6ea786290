"""Lock-step against free-running dense sweep kernel per shape (OSA_WS_FLOW=0/1, result-preserving knob):
one annealing run per shape on a bench-like schedule, ms of the sweep kernel."""
import json, os, sys
sys.path.insert(0, ".")
import numpy as np
from onesolver_b200 import Problem, capi
from onesolver_b200 import problems as gen

shapes = [(3072, "f32", 12), (5120, "f32", 8), (6144, "f32", 8), (8192, "f32", 4),
          (1536, "f64", 12), (2048, "f64", 8), (2560, "f64", 6), (3072, "f64", 6), (4096, "f64", 4),
          (2048, "f32", 16), (1024, "f64", 16)]
sweeps = 16
for n, prec, r in shapes:
    q = gen.dense_uniform_qubo(n, seed=100 + n)
    s = np.sqrt(n)
    sched = (0.02 * s) * (15.0) ** (np.arange(sweeps) / (sweeps - 1.0))
    tries = 148 * r * 2
    out = {"n": n, "prec": prec, "R": r, "tries": tries}
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32 if prec == "f32" else capi.SWEEP_F64) as p:
        for flow in ("0", "1"):
            os.environ["OSA_WS_FLOW"] = flow
            best = 1e30
            for rep in range(2):
                res = p.anneal(sched, sweeps, tries, mode=capi.MODE_SEQUENTIAL_SWEEP)
                best = min(best, res.stats["ms_sweep"])
            out["flow" + flow + "_ms"] = round(best, 2)
            out["R_seen"] = res.stats["traj_per_batch"]
    out["flow_speedup"] = round(out["flow0_ms"] / out["flow1_ms"], 3)
    print(json.dumps(out), flush=True)
