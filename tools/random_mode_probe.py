"""Throughput of the reference-faithful random-site mode (k_dense_generic / k_sparse) on the GPU."""
import os, sys, json
sys.path.insert(0, ".")
import numpy as np
from onesolver_b200 import Problem, capi
from onesolver_b200 import problems as gen

def geo(n, lo, hi):
    return lo * (hi / lo) ** (np.arange(n) / max(1, n - 1))

cases = ((4096, 8192, 8192, capi.SWEEP_F32), (1024, 16384, 8192, capi.SWEEP_F64), (128, 65536, 4096, capi.SWEEP_F64))
if os.environ.get("TRIES"):  # N = 4096 only, with a trajectory count that fills whole waves
    cases = ((4096, int(os.environ["TRIES"]), int(os.environ.get("ITERS", 2048)), capi.SWEEP_F32),)
if os.environ.get("CASES"):  # "n:tries:iters:f32|f64,..."
    cases = tuple((int(a), int(b), int(c), capi.SWEEP_F32 if d == "f32" else capi.SWEEP_F64)
                  for a, b, c, d in (x.split(":") for x in os.environ["CASES"].split(",")))
for n, tries, iters, prec in cases:
    q = gen.dense_uniform_qubo(n, seed=2024)
    s = np.sqrt(n)
    with Problem.dense(q, sweep_precision=prec) as p:
        for rep in range(2):
            r = p.anneal(geo(iters, 0.3 * s, 0.02 * s), iters, tries, mode=capi.MODE_RANDOM_SITE)
        st = r.stats
    print(json.dumps({"probe": "random_site_dense", "n": n, "tries": tries, "num_iter": iters,
                      "kernel": st["kernel_id"], "ms_sweep": round(st["ms_sweep"], 2),
                      "attempts_per_s": st["attempts"] / (st["ms_sweep"] * 1e-3),
                      "accept_frac": st["accepts"] / st["attempts"],
                      "row_gbs": st["accepts"] * n * (4 if prec == capi.SWEEP_F32 else 8) / (st["ms_sweep"] * 1e-3) / 1e9}))
