"""Parallel tempering: osa_pt_anneal against oracle/pt.py.

There is no reference implementation (the reference only recommends the method,
benchmarks/annealing/performance.md:54-59), so parity is against the CPU restatement of the
engine's own definition: bit-exact on instances with exactly representable coefficients (every
energy is then exact in any summation order), and through properties on float instances."""
import json
import os
import subprocess

import numpy as np
import pytest

from onesolver_b200 import Problem, capi, unpack_states
from onesolver_b200 import problems as gen

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def oracle_pt(*args, **kw):
    from oracle import pt
    return pt.parallel_tempering(*args, **kw)


@pytest.mark.parametrize("n,prec,dtype,rule", [
    (40, capi.SWEEP_F32, np.float32, capi.ACCEPT_BOLTZMANN),
    (33, capi.SWEEP_F64, np.float64, capi.ACCEPT_BOLTZMANN),
    (70, capi.SWEEP_F32, np.float32, capi.ACCEPT_REFERENCE),
])
def test_pt_matches_the_oracle_bit_for_bit(n, prec, dtype, rule):
    q = gen.dense_integer_qubo(n, seed=11 + n)
    betas = np.geomspace(0.05, 2.0, 6) if rule == capi.ACCEPT_BOLTZMANN else np.geomspace(0.5, 20.0, 6)[::-1].copy()
    if rule == capi.ACCEPT_REFERENCE:
        betas = np.sort(betas)
    groups, rounds, sweeps = 5, 9, 3
    with Problem.dense(q, sweep_precision=prec) as p:
        r = p.parallel_tempering(betas, groups, rounds, sweeps, accept_rule=rule,
                                 want_energies=True, want_states=True)
    o = oracle_pt(q, betas, groups, rounds, sweeps, accept_rule=rule, dtype=dtype)
    assert (r.best_states_packed == o["best_states"]).all()
    assert (r.best_energies == o["best_energies"]).all()
    assert r.energy == o["energy"] and r.index == o["index"]
    assert (r.state == o["state"]).all()
    assert r.stats["pt_swaps"] == o["swaps"] and o["swaps"] > 0
    assert r.stats["attempts"] == rounds * sweeps * n * groups * len(betas)


def test_pt_group_offset_and_prefix_stability():
    """Groups are keyed by their global id: a shard [first_group, ...) reproduces the matching
    slice of the full run, like first_try does for plain annealing."""
    q = gen.dense_integer_qubo(48, seed=5)
    betas = np.geomspace(0.05, 3.0, 4)
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as p:
        full = p.parallel_tempering(betas, 6, 6, 2, want_energies=True, want_states=True)
        part = p.parallel_tempering(betas, 3, 6, 2, first_group=2, want_energies=True, want_states=True)
    m = len(betas)
    assert (part.best_energies == full.best_energies[2 * m:5 * m]).all()
    assert (part.best_states_packed == full.best_states_packed[2 * m:5 * m]).all()


def test_pt_float_instance_properties():
    """U(-1,1) coefficients: returned energies are the exact energies of the returned states, the
    winner is the first minimum, and one replica per ladder is no worse than plain annealing at
    the coldest temperature started from the same spins."""
    n = 200
    q = gen.dense_uniform_qubo(n, seed=9)
    betas = np.geomspace(0.2, 8.0, 8)
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as p:
        r = p.parallel_tempering(betas, 16, 20, 4, want_energies=True, want_states=True)
        e = p.energy_batch(r.best_states_packed)
    assert np.allclose(e, r.best_energies, rtol=1e-12, atol=1e-9)
    x = unpack_states(r.best_states_packed, n).astype(np.float64)
    ref = np.einsum("ti,ij,tj->t", x, np.triu(q), x)
    assert np.allclose(ref, r.best_energies, rtol=1e-9, atol=1e-9)
    assert r.index == int(np.argmin(r.best_energies)) and r.energy == r.best_energies[r.index]
    assert r.stats["pt_swaps"] > 0


def test_pt_argument_checks():
    q = gen.dense_integer_qubo(16, seed=1)
    with Problem.dense(q) as p:
        with pytest.raises(capi.OsaError, match="strictly increasing"):
            p.parallel_tempering([1.0, 0.5], 1, 1, 1)
        with pytest.raises(capi.OsaError, match="num_rounds"):
            p.parallel_tempering([0.5, 1.0], 1, 0, 1)
    rowptr, col, val, diag = gen.sparse_random_graph(64, 4, seed=2)
    with Problem.csr(rowptr, col, val, diag) as p:
        with pytest.raises(capi.OsaError, match="dense problems"):
            p.parallel_tempering([0.5, 1.0], 1, 1, 1)


def test_pt_cli_finds_the_chimera128_ground_state(tmp_path):
    """one-solver-anneal --algorithm pt on the reference's own benchmark instance."""
    assert subprocess.run(["make", "-C", os.path.join(ROOT, "app")], capture_output=True).returncode == 0
    d = os.path.join(HERE, "golden", "chimera128")
    out = tmp_path / "pt.csv"
    r = subprocess.run([os.path.join(ROOT, "build", "bin", "one-solver-anneal"), "--input",
                        os.path.join(d, "001.qubo"), "--output", str(out), "--device-type", "gpu",
                        "--algorithm", "pt", "--accept", "boltzmann", "--num-replicas", "12",
                        "--num-iter", "200", "--sweeps-per-beta", "5", "--num-tries", "128",
                        "--beta-min", "0.1", "--beta-max", "10", "--stats"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Replica exchanges accepted" in r.stdout
    energy = float(out.read_text().splitlines()[1].split(",")[-1])
    assert abs(energy - (-235.867)) < 2e-3
