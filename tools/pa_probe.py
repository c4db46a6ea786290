"""Throughput of the samplers built on the resumable dense sweep kernel (population annealing, parallel
tempering) next to plain annealing with the same number of sweeps: how much the per-step round trips
(exact energies, resampling / exchanges, one kernel launch per step) cost."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
from onesolver_b200 import Problem, capi
from onesolver_b200 import problems as gen

for n, prec in ((1024, capi.SWEEP_F32), (4096, capi.SWEEP_F32)):
    q = gen.dense_uniform_qubo(n, seed=2024)
    s = np.sqrt(n)
    steps, sweeps = 32, 2
    betas = np.geomspace(0.02 * s, 0.3 * s, steps)  # the reference's rule: beta is a temperature
    betas = betas[::-1].copy()                      # ... so an annealing run walks it downwards
    tries = 148 * 16 * 4 if n == 1024 else 148 * 12 * 4
    with Problem.dense(q, sweep_precision=prec) as p:
        for rep in range(2):
            pa = p.population_annealing(betas, 4, tries // 4, sweeps, accept_rule=capi.ACCEPT_REFERENCE)
            sa = p.anneal(betas, steps, tries, sweeps_per_beta=sweeps, mode=capi.MODE_SEQUENTIAL_SWEEP)
            pt = p.parallel_tempering(np.sort(betas), tries // steps, steps, sweeps,
                                      accept_rule=capi.ACCEPT_REFERENCE)
    for name, r in (("plain annealing", sa), ("population annealing", pa), ("parallel tempering", pt)):
        st = r.stats
        print(json.dumps({"probe": name, "n": n, "trajectories": tries, "sweeps": steps * sweeps,
                          "ms_total": round(st["ms_total"], 2), "ms_sweep_phase": round(st["ms_sweep"], 2),
                          "attempts_per_s": st["attempts"] / (st["ms_total"] * 1e-3),
                          "launches": st["launches"], "exchanged_or_resampled": st["pt_swaps"],
                          "best_energy": r.energy}))
