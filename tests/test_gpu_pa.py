"""Population annealing: osa_pa_anneal against oracle/pa.py.

There is no reference implementation (the reference only recommends the method,
benchmarks/annealing/performance.md:54-59), so parity is against the CPU restatement of the
engine's own definition: bit-exact on instances with exactly representable coefficients (every
energy is then exact in any summation order, and the resampling works on integer weights, so no
reduction order enters), and through properties on float instances."""
import numpy as np
import pytest

from onesolver_b200 import Problem, capi, unpack_states
from onesolver_b200 import problems as gen

pytestmark = pytest.mark.gpu


def oracle_pa(*args, **kw):
    from oracle import pa
    return pa.population_annealing(*args, **kw)


@pytest.mark.parametrize("n,prec,dtype,rule,pops,size", [
    (40, capi.SWEEP_F32, np.float32, capi.ACCEPT_BOLTZMANN, 3, 50),
    (33, capi.SWEEP_F64, np.float64, capi.ACCEPT_BOLTZMANN, 2, 1500),   # more than one scan chunk
    (70, capi.SWEEP_F32, np.float32, capi.ACCEPT_REFERENCE, 4, 33),
    (24, capi.SWEEP_F64, np.float64, capi.ACCEPT_BOLTZMANN, 5, 1),      # populations of one
])
def test_pa_matches_the_oracle_bit_for_bit(gpu, n, prec, dtype, rule, pops, size):
    q = gen.dense_integer_qubo(n, seed=17 + n)
    # the reference's rule accepts with exp(-dE / beta): its "beta" is a temperature, so an
    # annealing run walks it downwards; the Boltzmann rule walks beta upwards
    betas = np.geomspace(0.03, 1.5, 7) if rule == capi.ACCEPT_BOLTZMANN else np.geomspace(30.0, 0.6, 7)
    sweeps = 2
    with Problem.dense(q, sweep_precision=prec) as p:
        r = p.population_annealing(betas, pops, size, sweeps, accept_rule=rule,
                                   want_energies=True, want_states=True)
    o = oracle_pa(q, betas, pops, size, sweeps, accept_rule=rule, dtype=dtype)
    assert (r.best_states_packed == o["best_states"]).all()
    assert (r.best_energies == o["best_energies"]).all()
    assert r.energy == o["energy"] and r.index == o["index"]
    assert (r.state == o["state"]).all()
    assert r.stats["pt_swaps"] == o["resampled"]
    assert (o["resampled"] > 0) == (size > 1)
    assert r.stats["attempts"] == len(betas) * sweeps * n * pops * size


def test_pa_population_offset(gpu):
    """Populations are keyed by their global id: a shard [first_population, ...) reproduces the
    matching slice of the full run, like first_try does for plain annealing."""
    q = gen.dense_integer_qubo(48, seed=5)
    betas = np.linspace(0.05, 2.0, 6)
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as p:
        full = p.population_annealing(betas, 6, 40, 2, want_energies=True, want_states=True)
        part = p.population_annealing(betas, 3, 40, 2, first_population=2, want_energies=True,
                                      want_states=True)
    assert (part.best_energies == full.best_energies[80:200]).all()
    assert (part.best_states_packed == full.best_states_packed[80:200]).all()


def test_pa_float_instance_properties_and_quality(gpu):
    """U(-1,1) coefficients, N = 200: returned energies are the exact energies of the returned
    states, the winner is the first minimum, resampling happened, and the population's best is at
    least as good as the best of as many independent annealing runs on the same schedule."""
    n = 200
    q = gen.dense_uniform_qubo(n, seed=9)
    betas = np.geomspace(0.2, 8.0, 24)
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as p:
        r = p.population_annealing(betas, 4, 512, 2, want_energies=True, want_states=True)
        e = p.energy_batch(r.best_states_packed)
        plain = p.anneal(betas, len(betas), 2048, sweeps_per_beta=2,
                         mode=capi.MODE_SEQUENTIAL_SWEEP, accept_rule=capi.ACCEPT_BOLTZMANN)
    assert np.allclose(e, r.best_energies, rtol=1e-12, atol=1e-9)
    x = unpack_states(r.best_states_packed, n).astype(np.float64)
    ref = np.einsum("ti,ij,tj->t", x, np.triu(q), x)
    assert np.allclose(ref, r.best_energies, rtol=1e-9, atol=1e-9)
    assert r.index == int(np.argmin(r.best_energies)) and r.energy == r.best_energies[r.index]
    assert r.stats["pt_swaps"] > 0
    assert r.energy <= plain.energy + 0.02 * abs(plain.energy)


def test_pa_argument_checks(gpu):
    q = gen.dense_integer_qubo(16, seed=1)
    with Problem.dense(q) as p:
        with pytest.raises(capi.OsaError, match="population_size"):
            p.population_annealing([0.5, 1.0], 1, 0, 1)
        with pytest.raises(capi.OsaError, match="positive finite"):
            p.population_annealing([0.5, -1.0], 1, 4, 1)
        with pytest.raises(capi.OsaError, match="sweeps_per_step"):
            p.population_annealing([0.5, 1.0], 1, 4, 0)
    rowptr, col, val, diag = gen.sparse_random_graph(64, 4, seed=2)
    with Problem.csr(rowptr, col, val, diag) as p:
        with pytest.raises(capi.OsaError, match="dense problems"):
            p.population_annealing([0.5, 1.0], 1, 4, 1)


def test_pa_cli_finds_the_chimera128_ground_state(gpu, tmp_path):
    """one-solver-anneal --algorithm pa on the reference's own benchmark instance (the reference's
    annealer never found this ground state, benchmarks/annealing/performance.md:39-45)."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    assert subprocess.run(["make", "-C", os.path.join(root, "app")], capture_output=True).returncode == 0
    out = tmp_path / "pa.csv"
    r = subprocess.run([os.path.join(root, "build", "bin", "one-solver-anneal"), "--input",
                        os.path.join(here, "golden", "chimera128", "001.qubo"), "--output", str(out),
                        "--device-type", "gpu", "--algorithm", "pa", "--accept", "boltzmann",
                        "--num-replicas", "1024", "--num-iter", "100", "--sweeps-per-beta", "4",
                        "--num-tries", "4", "--schedule-type", "linear", "--beta-min", "0.1",
                        "--beta-max", "10", "--stats"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "Replicas resampled" in r.stdout
    energy = float(out.read_text().splitlines()[1].split(",")[-1])
    assert abs(energy - (-235.867)) < 2e-3
