"""CPU restatement of osa_pa_anneal (population annealing around the dense sweep kernel).

TEST INFRASTRUCTURE: imported by tests/ only.  The reference has no population annealing (its
report names it first among the samplers it recommends, benchmarks/annealing/performance.md:54-59),
so this restates the engine's own definition (include/onesolver_b200.h, osa_pa_anneal) step by step
with the oracle's primitives: the bit-exact sweep replay, the reference energy formula, and the
resampling step orc_pa_resample (integer weights, systematic resampling).  Energies are exact only
up to summation order, so bit-exact agreement with the GPU is asserted on instances with exactly
representable coefficients.  (The engine carries the local fields of a replica from round to
round; this restatement rebuilds them from the spins every round.  The two are the same walk whenever
every partial sum is exact, which is the case the bit-exact tests use.)
"""
import numpy as np

from . import binding as ob
from .pt import initial_states


def population_annealing(qsym, betas, num_populations, population_size, sweeps_per_step, seed=1234,
                         first_population=0, accept_rule=1, dtype=np.float64):
    """-> dict(best_energies[tries], best_states[tries][nw], energy, index, state, resampled)."""
    q = np.asarray(qsym, dtype=np.float64)
    n = q.shape[0]
    betas = np.asarray(betas, dtype=np.float64)
    m = population_size
    tries = num_populations * m
    first_try = first_population * m
    qoff, diag = ob.split_dense(q, dtype)
    inv_t = (1.0 / betas) if accept_rule == 0 else betas

    cur = initial_states(seed, first_try, tries, n)
    e_cur = ob.energy_packed(q, cur)
    best_e = np.full(tries, np.inf)
    keep = np.zeros_like(cur)
    resampled = 0
    for step in range(len(betas)):
        ts = betas[step] if accept_rule == 0 else 1.0 / betas[step]
        ts_traj = np.full(tries, ts).astype(dtype)
        best_rel, step_best, final = ob.replay_dense_round(
            qoff, diag, ts_traj, cur, sweeps_per_step, seed, first_try, step * sweeps_per_step)
        cand = e_cur + best_rel
        better = cand < best_e
        keep[better] = step_best[better]
        best_e[better] = cand[better]
        cur = final
        if step + 1 == len(betas):
            break
        e_cur = ob.energy_packed(q, cur)
        neg_db = -(inv_t[step + 1] - inv_t[step])
        nxt, e_nxt = cur.copy(), e_cur.copy()
        for g in range(num_populations):
            base = g * m
            src = ob.pa_resample(e_cur[base:base + m], neg_db, seed, first_population + g, step)
            nxt[base:base + m] = cur[base + src]
            e_nxt[base:base + m] = e_cur[base + src]
            resampled += int((src != np.arange(m)).sum())
        cur, e_cur = nxt, e_nxt
    energies = ob.energy_packed(q, keep)
    index = int(np.argmin(energies))  # first minimum, like std::min_element
    nbits = ((keep[index][:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).reshape(-1)[:n]
    return {"best_energies": energies, "best_states": keep, "energy": float(energies[index]),
            "index": first_try + index, "state": nbits.astype(np.uint8), "resampled": resampled}
