"""CPU checks of the population-annealing restatement (oracle/pa.py, orc_pa_resample,
orc_det_exp) that the GPU tests compare against."""
import math

import numpy as np

from oracle import binding as ob
from oracle import pa
from onesolver_b200 import problems as gen


def test_det_exp_is_exp_to_an_ulp_on_the_range_used():
    rng = np.random.default_rng(5)
    xs = np.concatenate([[0.0, -1e-300, -1e-9, -0.5, -1.0, -59.999], -rng.uniform(0, 60, 2000)])
    for x in xs:
        got, want = ob.det_exp(x), math.exp(x)
        assert abs(got - want) <= 4.5e-16 * want, (x, got, want)
    assert ob.det_exp(0.0) == 1.0
    assert ob.det_exp(-60.5) == 0.0 and ob.det_exp(-1e9) == 0.0  # below the 40-bit weights
    assert ob.det_exp(float("nan")) == 0.0


def test_resampling_is_systematic_and_keeps_the_population_size():
    """Slot j continues from a replica whose cumulative weight brackets (j + u) W / M: the number
    of copies of replica i is floor or ceil of its expected count M w_i / W, the sources come in
    ascending order, and equal weights leave the population as it is."""
    rng = np.random.default_rng(11)
    m = 257
    e = rng.normal(0.0, 3.0, m)
    for db in (0.05, 0.7, -0.3):
        src = ob.pa_resample(e, -db, 1234, 9, 4)
        assert src.shape == (m,) and (np.diff(src) >= 0).all() and 0 <= src.min() and src.max() < m
        w = np.exp(-db * (e - e.min() if db > 0 else e - e.max()))
        expect = m * w / w.sum()
        copies = np.bincount(src, minlength=m)
        assert (copies >= np.floor(expect - 1e-6)).all() and (copies <= np.ceil(expect + 1e-6)).all()
    same = ob.pa_resample(np.full(m, -2.5), -0.4, 1234, 0, 0)
    assert (same == np.arange(m)).all()
    # a different step or population draws another offset; the same key the same one
    a = ob.pa_resample(e, -0.7, 1234, 9, 4)
    assert (a == ob.pa_resample(e, -0.7, 1234, 9, 4)).all()
    assert any((a != ob.pa_resample(e, -0.7, 1234, 9 + k, 4)).any() for k in range(1, 6))
    # a replica far above the rest (weight below 2^-40) dies out, the best one multiplies
    e2 = np.zeros(64)
    e2[7], e2[9] = 200.0, -3.0
    src = ob.pa_resample(e2, -1.0, 1234, 0, 0)
    assert 7 not in src and (src == 9).sum() >= 10


def test_one_temperature_is_plain_annealing_at_fixed_beta():
    """A schedule with a single temperature never resamples: every replica is a sequential-sweep
    run at constant beta, i.e. the pinned replay with a flat schedule."""
    n = 36
    q = gen.dense_integer_qubo(n, seed=2)
    r = pa.population_annealing(q, [0.7], 2, 9, 5, accept_rule=1, dtype=np.float64)
    assert r["resampled"] == 0
    _, best, _, _ = ob.replay_dense(q, np.full(5, 0.7), 5, 18, mode=1, accept_rule=1,
                                    dtype=np.float64)
    assert (best == r["best_states"]).all()
    assert (ob.energy_packed(q, best) == r["best_energies"]).all()


def test_population_annealing_finds_the_ground_state_of_a_small_instance():
    n = 20
    q = gen.dense_integer_qubo(n, seed=31)
    _, gs = ob.ref_exhaustive(q, n, 4)
    betas = np.linspace(0.02, 1.5, 12)
    r = pa.population_annealing(q, betas, 2, 48, 2, accept_rule=1, dtype=np.float32)
    assert r["resampled"] > 0
    assert r["energy"] == gs
    x = ((r["best_states"][:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(96, -1)[:, :n]
    for t in (0, 17, 95):
        assert ob.ref_energy(q, x[t].astype(np.int8)) == r["best_energies"][t]
    assert r["index"] == int(np.argmin(r["best_energies"]))
    # populations are keyed by their global id: population 1 of this run == population 0 of a run
    # that starts at first_population = 1
    r1 = pa.population_annealing(q, betas, 1, 48, 2, accept_rule=1, dtype=np.float32,
                                 first_population=1)
    assert (r1["best_states"] == r["best_states"][48:]).all()
