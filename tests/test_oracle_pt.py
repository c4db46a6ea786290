"""CPU checks of the parallel-tempering restatement (oracle/pt.py) that the GPU tests compare
against: its pieces must agree with the already pinned oracle primitives."""
import numpy as np

from oracle import binding as ob
from oracle import pt
from onesolver_b200 import problems as gen


def test_initial_states_are_the_engine_stream():
    """Round 0 starts from the spins the sweep kernels draw themselves (STREAM_INIT)."""
    n, tries = 70, 5
    states = pt.initial_states(1234, 3, tries, n)
    for t in range(tries):
        for j in range(n):
            assert ((int(states[t, j >> 5]) >> (j & 31)) & 1) == ob.load().orc_init_bit(1234, 3 + t, j)
    assert (states[:, 2] >> (n - 64)).max() == 0  # padding bits are clear


def test_one_rung_without_exchange_is_plain_annealing_at_fixed_beta():
    """A ladder with a single temperature never exchanges: every replica is a sequential-sweep
    annealing run at constant beta, i.e. the pinned replay with a flat schedule."""
    n = 36
    q = gen.dense_integer_qubo(n, seed=2)
    rounds, sweeps, groups = 4, 3, 6
    r = pt.parallel_tempering(q, [0.7], groups, rounds, sweeps, accept_rule=1, dtype=np.float64)
    assert r["swaps"] == 0
    sched = np.full(rounds * sweeps, 0.7)
    _, best, _, _ = ob.replay_dense(q, sched, rounds * sweeps, groups, mode=1, accept_rule=1,
                                    dtype=np.float64)
    assert (ob.energy_packed(q, best) == r["best_energies"]).all()
    assert (best == r["best_states"]).all()


def test_ladder_stays_a_permutation_and_energies_are_exact():
    n = 40
    q = gen.dense_integer_qubo(n, seed=7)
    betas = np.geomspace(0.05, 2.0, 5)
    r = pt.parallel_tempering(q, betas, 3, 12, 2, accept_rule=1, dtype=np.float32)
    assert r["swaps"] > 0
    for g in range(3):
        assert sorted(r["temp_of_slot"][g * 5:(g + 1) * 5]) == list(range(5))
    x = ((r["best_states"][:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(15, -1)[:, :n]
    for t in range(15):
        assert ob.ref_energy(q, x[t].astype(np.int8)) == r["best_energies"][t]
    assert r["energy"] == r["best_energies"].min() and r["index"] == int(np.argmin(r["best_energies"]))
