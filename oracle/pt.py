"""CPU restatement of osa_pt_anneal (parallel tempering around the dense sweep kernel).

TEST INFRASTRUCTURE: imported by tests/ only.  The reference has no parallel tempering (its report
recommends one, benchmarks/annealing/performance.md:54-59), so this restates the engine's own
definition (include/onesolver_b200.h, osa_pt_anneal; onesolver_b200/csrc/osa_pt.cu) step by step
with the oracle's primitives: the bit-exact sweep replay, the reference energy formula, Philox and
the deterministic -ln(u).  Energies are exact only up to summation order, so bit-exact agreement
with the GPU is asserted on instances with exactly representable coefficients.  (The engine carries the local fields of a replica from round to
round; this restatement rebuilds them from the spins every round.  The two are the same walk whenever
every partial sum is exact, which is the case the bit-exact tests use.)
"""
import numpy as np

from . import binding as ob

STREAM_INIT, STREAM_PT = 0, 3


def initial_states(seed, first_try, tries, n):
    nw = (n + 31) // 32
    out = np.zeros((tries, nw), dtype=np.uint32)
    for t in range(tries):
        for k in range(nw):
            word = int(ob.engine_draw(seed, first_try + t, STREAM_INIT, k >> 2, 0)[k & 3])
            valid = n - 32 * k
            if valid < 32:
                word &= (1 << valid) - 1
            out[t, k] = word
    return out


def parallel_tempering(qsym, betas, num_groups, num_rounds, sweeps_per_round, seed=1234,
                       first_group=0, accept_rule=1, dtype=np.float64):
    """-> dict(best_energies[tries], best_states[tries][nw], energy, index, state, swaps)."""
    q = np.asarray(qsym, dtype=np.float64)
    n = q.shape[0]
    betas = np.asarray(betas, dtype=np.float64)
    m = betas.shape[0]
    tries = num_groups * m
    first_try = first_group * m
    qoff, diag = ob.split_dense(q, dtype)
    ts_rung = (betas if accept_rule == 0 else 1.0 / betas).astype(dtype)
    inv_t = (1.0 / betas) if accept_rule == 0 else betas
    dinv = inv_t[:-1] - inv_t[1:]

    cur = initial_states(seed, first_try, tries, n)
    e_cur = ob.energy_packed(q, cur)
    temp_of_slot = np.tile(np.arange(m, dtype=np.int64), num_groups)
    slot_of_temp = temp_of_slot.copy()
    ts_traj = ts_rung[temp_of_slot].copy()
    best_e = np.full(tries, np.inf)
    keep = np.zeros_like(cur)
    swaps = 0
    for rnd in range(num_rounds):
        best_rel, round_best, final = ob.replay_dense_round(
            qoff, diag, ts_traj, cur, sweeps_per_round, seed, first_try, rnd * sweeps_per_round)
        cand = e_cur + best_rel
        better = cand < best_e
        keep[better] = round_best[better]
        best_e[better] = cand[better]
        cur = final
        e_cur = ob.energy_packed(q, cur)
        for g in range(num_groups):
            base = g * m
            for j in range(rnd & 1, m - 1, 2):
                a, b = int(slot_of_temp[base + j]), int(slot_of_temp[base + j + 1])
                x = dinv[j] * (e_cur[base + a] - e_cur[base + b])
                accept = x >= 0.0
                if not accept:
                    w = int(ob.engine_draw(seed, first_group + g, STREAM_PT, j, rnd)[0])
                    accept = -x < float(np.float32(ob.neglogf(w)))
                if accept:
                    slot_of_temp[base + j], slot_of_temp[base + j + 1] = b, a
                    temp_of_slot[base + a], temp_of_slot[base + b] = j + 1, j
                    ts_traj[base + a], ts_traj[base + b] = ts_rung[j + 1], ts_rung[j]
                    swaps += 1
    energies = ob.energy_packed(q, keep)
    index = int(np.argmin(energies))  # first minimum, like std::min_element
    nbits = ((keep[index][:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).reshape(-1)[:n]
    return {"best_energies": energies, "best_states": keep, "energy": float(energies[index]),
            "index": first_try + index, "state": nbits.astype(np.uint8), "swaps": swaps,
            "temp_of_slot": temp_of_slot}
