"""CPU tests: pin the oracle (oracle/osa_oracle.c, oracle/qubo_format.py) against the reference's
own golden vectors (tests/golden/*, generated from /root/reference by make_golden.py)."""
import json
import math
import os

import numpy as np
import pytest

from oracle import binding as ob
from oracle.qubo_format import QuboFormatError, ising_to_qubo, load_qubo, parse_qubo

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
G = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))


def dense_from(n, lin, quad):
    return ob.ref_flatten(n, lin, quad).reshape(n, n)


def test_philox_known_answers():
    """Philox4x32-10 KATs (Random123 kat_vectors)."""
    kat = [([0, 0, 0, 0], [0, 0], "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
           ([0xffffffff] * 4, [0xffffffff] * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            "d16cfe09 94fdcceb 5001e420 24126ea1")]
    for ctr, key, want in kat:
        assert " ".join("%08x" % v for v in ob.philox(ctr, key)) == want


def test_neglog_is_positive_monotone_and_accurate():
    ws = np.unique(np.concatenate([np.random.default_rng(0).integers(0, 2 ** 32, 5000),
                                   [0, 1, 2, 2 ** 31, 2 ** 32 - 2, 2 ** 32 - 1]]))
    vals = np.array([ob.neglogf(int(w)) for w in ws])
    assert (vals > 0).all()
    assert (np.diff(vals) <= 1e-6).all()  # non-increasing in w up to float rounding
    exact = -np.log((2.0 * ws + 1.0) / 2.0 ** 33)
    inner = ws < 2 ** 32 - 2 ** 12
    assert np.max(np.abs(vals[inner] - exact[inner]) / exact[inner]) < 5e-3
    assert np.max(np.abs(vals - exact)) < 1e-6 * 23 + 6e-8


def test_flatten_matches_reference_layout():
    f = G["flatten"]
    lin = {int(k): v for k, v in f["linear"].items()}
    quad = {(i, j): v for i, j, v in f["quadratic"]}
    assert ob.ref_flatten(f["n"], lin, quad).tolist() == f["expected"]


def test_exhaustive_ground_energies_of_reference_tests():
    for case in G["exhaustive_test"]:
        n, lin, quad = parse_qubo(case["qubo"])
        q = dense_from(n, lin, quad)
        for ranges in (1, 3, 12):
            state, e = ob.ref_exhaustive(q, n, ranges)
            assert abs(e - case["energy"]) < 1e-13
            assert abs(ob.ref_energy(q, state) - case["energy"]) < 1e-13


def test_shipped_examples_ground_states():
    for name, want in G["examples_ground"].items():
        n, lin, quad = load_qubo(os.path.join(ROOT, "examples", name))
        q = dense_from(n, lin, quad)
        state, e = ob.ref_exhaustive(q, n, 4)
        assert abs(e - want["energy"]) < 1e-12 and state.tolist() == want["state"]
    with pytest.raises(QuboFormatError):  # "p  qubo" header: the reference rejects this file
        load_qubo(os.path.join(ROOT, "examples", "dwave_doc.qubo"))


def test_parser_accept_reject_set_of_io_test():
    n, lin, quad = parse_qubo(G["io_valid"])
    assert (n, lin, quad) == (3, {0: -0.5, 2: -0.7}, {(0, 1): 2.0, (1, 2): 4.0})
    for text in G["io_malformed"] + [G["io_lower_triangle"], G["io_count_mismatch"]]:
        with pytest.raises(QuboFormatError):
            parse_qubo(text)


def test_solution_csv_format():
    s = G["io_solution_csv"]
    assert ob.ref_solution_csv(s["state"], s["energy"]) == s["text"]


def test_schedules_follow_reference_quirks():
    lin = ob.ref_schedule("linear", 0.1, 1.0, 100)
    geo = ob.ref_schedule("geometric", 0.1, 1.0, 100)
    assert lin[0] == 0.1 and lin[-1] == 0.1 + 1.0            # ends at beta_min + beta_max
    assert lin[37] == 0.1 + 1.0 * 37 / 99.0
    alpha = math.pow(1.0 / 0.1, 1.0 / 99)
    ref = [0.1]
    for _ in range(99):
        ref.append(ref[-1] * alpha)                             # iterated product, not pow(alpha, i)
    assert geo.tolist() == ref


@pytest.mark.parametrize("size", ["128", "512"])
def test_chimera_results_of_the_reference_cli(size):
    """(state, energy) rows the reference CLI wrote (benchmarks/annealing/results): the oracle's
    parser + flatten + energy must reproduce every energy to the CSV's 6 significant digits, and
    no annealing result may lie below the tensor-network ground state."""
    d = os.path.join(HERE, "golden", f"chimera{size}")
    n, lin, quad = load_qubo(os.path.join(d, "001.qubo"))
    assert n == int(size)
    q = dense_from(n, lin, quad)
    rows = json.load(open(os.path.join(d, "reference_results_sample.json")))
    assert len(rows) >= 40
    for r in rows:
        state = np.array([int(c) for c in r["state"]], dtype=np.int8)
        e = ob.ref_energy(q, state)
        assert abs(e - r["energy"]) <= 5e-6 * abs(r["energy"]) + 1e-9
    # Ising original -> QUBO conversion (convert_qbsolv_to_coo.py:23-37) and the TN ground state
    n2, lin2, quad2, offset = ising_to_qubo(open(os.path.join(d, "001.ising.txt")).read())
    assert n2 == n and set(quad2) == set(quad)
    assert max(abs(lin[i] - lin2[i]) for i in lin) < 1e-9
    assert max(abs(quad[k] - quad2[k]) for k in quad) < 1e-9
    tn = json.load(open(os.path.join(d, "groundstate_TN.json")))
    x = (np.array(tn["spins"]) + 1) // 2
    e_tn = ob.ref_energy(q, x.astype(np.int8))
    assert abs(e_tn + offset - tn["ising_energy"]) < 1e-4
    assert all(r["energy"] >= e_tn - 1e-6 for r in rows)


def test_reference_restatement_and_replay_walk_the_same_trajectories():
    """With exactly representable coefficients the new algorithm (local fields, dE, threshold
    form of the acceptance test) must reproduce the reference loop (full energy recompute,
    exp((E_cur-E_new)/beta) > u) trajectory by trajectory."""
    from onesolver_b200 import problems as gen
    n = 20
    q = gen.dense_integer_qubo(n, seed=4)
    sched = ob.ref_schedule("geometric", 0.1, 10.0, 300)
    _, states, energies = ob.ref_anneal(q, n, sched, 300, 500)
    for dtype in (np.float64, np.float32):
        _, best, _, cnt = ob.replay_dense(q, sched, 300, 500, mode=0, dtype=dtype)
        assert cnt.attempts == 300 * 500
        assert (ob.energy_packed(q, best) == energies).all()
    # prefix stability / determinism: trajectory i does not depend on num_tries
    _, states_small, energies_small = ob.ref_anneal(q, n, sched, 300, 100)
    assert (states_small == states[:100]).all() and (energies_small == energies[:100]).all()


def test_sparse_and_dense_replays_agree_on_integer_instances():
    from onesolver_b200 import problems as gen
    rowptr, col, val, diag = gen.sparse_random_graph(60, 5, seed=2, integer=True)
    q = gen.csr_to_dense(rowptr, col, val, diag)
    sched = ob.ref_schedule("linear", 0.5, 5.0, 6)
    for mode in (0, 1):
        _, a, _, ca = ob.replay_csr(rowptr, col, val, diag, sched, 6, 40, mode=mode)
        _, b, _, cb = ob.replay_dense(q, sched, 6, 40, mode=mode)
        assert (a == b).all() and ca.accepts == cb.accepts
