"""GPU parity tests (run with -m gpu on the B200 box): the CUDA engine, called through
the C ABI, against the oracle (oracle/osa_oracle.c) on the same seeded inputs.

Bars: spin sequences / best states bit-exact vs the host replay; returned energies
equal to the reference energy function (annealing.hpp:31-40) to 1e-9 relative (fp64);
best state == exhaustive ground state for N <= 30.
"""
import numpy as np
import pytest

from onesolver_b200 import Problem, capi, pack_states, unpack_states
from onesolver_b200 import problems as gen
from oracle import binding as ob

pytestmark = pytest.mark.gpu

REL = 1e-9  # north_star tolerance for fp64 energies


def geo(num_iter, lo=0.1, hi=1.0):
    return ob.ref_schedule("geometric", lo, hi, num_iter)


def assert_states_equal(got, want, what):
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, f"{what}: {bad.size}/{got.shape[0]} trajectories differ, first {bad[:8]}"


def run_and_compare_dense(q, sched, num_iter, num_tries, mode, dtype, spb=1, kernel=capi.KID_AUTO,
                          accept_rule=capi.ACCEPT_REFERENCE, first_try=0, seed=1234):
    prec = capi.SWEEP_F32 if dtype == np.float32 else capi.SWEEP_F64
    with Problem.dense(q, sweep_precision=prec) as prob:
        res = prob.anneal(sched, num_iter, num_tries, sweeps_per_beta=spb, mode=mode,
                          accept_rule=accept_rule, kernel_variant=kernel, first_try=first_try,
                          seed=seed, want_energies=True, want_states=True, want_trace=True)
    r = res.stats["traj_per_batch"]
    with ob.trace(num_tries) as tr:
        best_rel, best, _, cnt = ob.replay_dense(q, sched, num_iter, num_tries, sweeps_per_beta=spb,
                                                 mode=mode, accept_rule=accept_rule, seed=seed,
                                                 first_try=first_try, dtype=dtype, batch_r=r)
    # the spin SEQUENCE of every trajectory (north_star criterion 2): hash of all accepted flips
    bad = np.nonzero(res.trace_hash != tr.hashes)[0]
    assert bad.size == 0, f"flip traces differ for {bad.size}/{num_tries} trajectories, first {bad[:8]}"
    assert_states_equal(res.best_states_packed, best, "best states vs host replay")
    assert res.stats["accepts"] == cnt.accepts
    assert res.stats["attempts"] == cnt.attempts
    assert res.stats["row_fetches"] == cnt.row_fetches
    e_ref = ob.energy_packed(q, best)
    np.testing.assert_allclose(res.best_energies, e_ref, rtol=REL, atol=1e-12)
    k = int(np.argmin(e_ref))  # first minimum, like std::min_element
    assert res.index == first_try + k
    assert abs(res.energy - e_ref[k]) <= REL * max(1.0, abs(e_ref[k]))
    np.testing.assert_array_equal(res.state, unpack_states(best[k], q.shape[0])[0])
    return res, cnt


# ---------------------------------------------------------------- config 1 / 2
def test_config1_test1_qubo_defaults(gpu):
    """BASELINE config 1 on the GPU: test1.qubo, 100 iters x 100 tries, geometric 0.1->1.0."""
    lin = {0: -5, 1: -3, 2: -8, 3: -6}
    quad = {(0, 1): 2, (0, 2): 4, (0, 3): 0, (1, 2): 1, (1, 3): 0, (2, 3): 5}
    q = ob.ref_flatten(4, lin, quad).reshape(4, 4)
    res, _ = run_and_compare_dense(q, geo(100), 100, 100, capi.MODE_RANDOM_SITE, np.float64)
    assert res.energy == -12.0
    np.testing.assert_array_equal(res.state, [1, 1, 0, 1])
    # trajectory-level agreement with the reference-faithful restatement (exact arithmetic)
    _, _, e_ref = ob.ref_anneal(q, 4, geo(100), 100, 100)
    assert (res.best_energies == e_ref).mean() >= 0.99


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_config2_n24_matches_exhaustive_and_reference(gpu, dtype):
    """BASELINE config 2: dense N=24, 4096 tries, best state == exhaustive ground state."""
    n = 24
    q = gen.dense_integer_qubo(n, seed=2026)
    sched = geo(400, 0.5, 20.0)
    res, _ = run_and_compare_dense(q, sched, 400, 4096, capi.MODE_RANDOM_SITE, dtype)
    gs_state, gs_energy = ob.ref_exhaustive(q, n, num_ranges=8)
    assert res.energy == gs_energy
    assert ob.ref_energy(q, res.state.astype(np.int8)) == gs_energy
    # ground-state hit probability vs the reference-faithful loop (same streams)
    _, _, e_ref = ob.ref_anneal(q, n, sched, 400, 4096)
    p_gpu = (res.best_energies == gs_energy).mean()
    p_ref = (e_ref == gs_energy).mean()
    ci = 3.0 * np.sqrt(max(p_ref * (1 - p_ref), 1e-4) / 4096) * np.sqrt(2)
    assert abs(p_gpu - p_ref) <= ci, (p_gpu, p_ref, ci)
    assert (res.best_energies == e_ref).mean() >= 0.98
    # ... and on INDEPENDENT streams (other seed, disjoint trajectory ids): the two samples share
    # nothing but the instance, so this is the statistical comparison north_star asks for.
    # 3 sigma of the difference of two independent binomial proportions.
    _, _, e_ind = ob.ref_anneal(q, n, sched, 400, 4096, seed=987654321, first_try=1 << 20)
    p_ind = (e_ind == gs_energy).mean()
    pm = 0.5 * (p_gpu + p_ind)
    ci_ind = 3.0 * np.sqrt(max(pm * (1 - pm), 1e-4) * 2.0 / 4096)
    assert abs(p_gpu - p_ind) <= ci_ind, (p_gpu, p_ind, ci_ind)


def test_n24_fractional_coefficients(gpu):
    n = 24
    q = gen.dense_uniform_qubo(n, seed=7)
    res, _ = run_and_compare_dense(q, geo(300, 0.05, 2.0), 300, 2048, capi.MODE_RANDOM_SITE,
                                   np.float64)
    _, gs_energy = ob.ref_exhaustive(q, n, num_ranges=5)
    assert abs(res.energy - gs_energy) <= REL * abs(gs_energy)


# ---------------------------------------------------------------- dense, sequential sweeps
@pytest.mark.parametrize("n,dtype,tries,sweeps", [
    (5, np.float64, 20, 6), (33, np.float32, 40, 5), (64, np.float64, 33, 4),
    (100, np.float32, 50, 4), (300, np.float64, 40, 3), (513, np.float64, 20, 2),
    (1024, np.float32, 24, 2), (1100, np.float32, 17, 2), (1024, np.float64, 16, 2),
    (2100, np.float32, 9, 1), (2048, np.float64, 9, 1), (4096, np.float32, 8, 1),
    # the remaining register/ring shapes of launch_dense_seq_ws (ld/1024 = 5..8, ld/512 = 5..8)
    (5000, np.float32, 9, 1), (6100, np.float32, 9, 1), (7000, np.float32, 5, 1),
    (8192, np.float32, 5, 1), (2500, np.float64, 7, 1), (3000, np.float64, 7, 1),
    (3500, np.float64, 5, 1), (4096, np.float64, 5, 1),
])
def test_dense_seq_bit_exact(gpu, n, dtype, tries, sweeps):
    q = gen.dense_uniform_qubo(n, seed=100 + n)
    scale = np.sqrt(n)
    sched = geo(sweeps, 0.02 * scale, 0.6 * scale)
    res, cnt = run_and_compare_dense(q, sched, sweeps, tries, capi.MODE_SEQUENTIAL_SWEEP, dtype)
    assert res.stats["kernel_id"] == capi.KID_DENSE_SEQ
    assert cnt.accepts > 0


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 63, 64, 65])
@pytest.mark.parametrize("mode", [capi.MODE_RANDOM_SITE, capi.MODE_SEQUENTIAL_SWEEP])
def test_tiny_and_block_boundary_sizes(gpu, n, mode):
    """N around the 32-site block / spin-word boundary (and N = 1), a number of tries that is not a
    multiple of the trajectories per CTA, and a single trajectory; integer coefficients, so fp32
    and fp64 fields give the reference energies exactly."""
    rng = np.random.default_rng(n)
    a = rng.integers(-5, 6, size=(n, n)).astype(np.float64)
    q = np.triu(a, 1)
    q = q + q.T + np.diag(np.diag(a))
    sched = ob.ref_schedule("linear", 0.5, 5.0, 6)
    for dtype in (np.float64, np.float32):
        for tries in (37, 1):
            res, _ = run_and_compare_dense(q, sched, 6, tries, mode, dtype)
            assert (res.best_energies == ob.energy_packed(q, res.best_states_packed)).all()


def test_sparse_tiny_sizes_and_isolated_sites(gpu):
    """CSR instances with fewer sites than a block, a site without neighbours and one trajectory."""
    cases = [gen.sparse_random_graph(n, deg, seed=900 + n, integer=True) + (tries,)
             for n, deg, tries in [(2, 1, 5), (7, 2, 33), (33, 3, 1), (65, 4, 40)]]
    # five sites on a path 0-1, 3-4 with site 2 isolated (empty CSR row), and a single site
    cases.append((np.array([0, 1, 2, 2, 3, 4], dtype=np.int64), np.array([1, 0, 4, 3], dtype=np.int32),
                  np.array([2.0, 2.0, -3.0, -3.0]), np.array([-1.0, 1.0, -2.0, 1.0, 1.0]), 9))
    cases.append((np.array([0, 0], dtype=np.int64), np.array([], dtype=np.int32), np.array([]),
                  np.array([-4.0]), 3))
    for rowptr, col, val, diag, tries in cases:
        n = len(diag)
        sched = ob.ref_schedule("linear", 0.5, 5.0, 5)
        for mode in (capi.MODE_SEQUENTIAL_SWEEP, capi.MODE_RANDOM_SITE):
            with Problem.csr(rowptr, col, val, diag, sweep_precision=capi.SWEEP_F64) as prob:
                res = prob.anneal(sched, 5, tries, mode=mode, want_energies=True, want_states=True)
            _, best, _, cnt = ob.replay_csr(rowptr, col, val, diag, sched, 5, tries, mode=mode,
                                            dtype=np.float64)
            assert_states_equal(res.best_states_packed, best, f"sparse n={n}")
            assert res.stats["accepts"] == cnt.accepts
            q = gen.csr_to_dense(rowptr, col, val, diag)
            assert (res.best_energies == ob.energy_packed(q, best)).all()


def test_dense_seq_sweeps_per_beta_and_boltzmann(gpu):
    q = gen.dense_uniform_qubo(200, seed=5)
    sched = ob.ref_schedule("linear", 0.05, 2.0, 4)
    run_and_compare_dense(q, sched, 4, 30, capi.MODE_SEQUENTIAL_SWEEP, np.float32, spb=3,
                          accept_rule=capi.ACCEPT_BOLTZMANN)


# n = 150: the trajectory builds its own initial field, short rows; 300: initial fields from the
# shared-fetch kernel (osa_dense_init.cu), short rows; 1100 / 2500: shared initial fields and the
# software-pipelined row add (one and several rounds, ragged N); 4100 in fp32: the rows of a
# batch's accepted flips are added in one pass over the fields while the walk follows them through
# gathered elements (ragged last piece)
@pytest.mark.parametrize("n", [150, 300, 1100, 2500, 4100])
@pytest.mark.parametrize("mode", [capi.MODE_RANDOM_SITE, capi.MODE_SEQUENTIAL_SWEEP])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_dense_generic_kernel_bit_exact(gpu, mode, dtype, n):
    q = gen.dense_uniform_qubo(n, seed=11)
    iters = 3 if mode == capi.MODE_SEQUENTIAL_SWEEP else 700
    sched = geo(iters, 0.3, 6.0)
    res, _ = run_and_compare_dense(q, sched, iters, 37, mode, dtype, kernel=capi.KID_DENSE_GENERIC)
    assert res.stats["kernel_id"] == capi.KID_DENSE_GENERIC


# rows of at least four rounds: both forms of the row add, whatever the library would choose
# (OSA_GEN_BATCH; the batched form is the default for fp32 rows of 16-20 KiB only)
@pytest.mark.parametrize("force", ["0", "1"])
@pytest.mark.parametrize("n,dtype", [(2500, np.float64), (4100, np.float32), (4100, np.float64),
                                     (6100, np.float32)])
def test_dense_generic_row_add_forms_bit_exact(gpu, monkeypatch, force, n, dtype):
    monkeypatch.setenv("OSA_GEN_BATCH", force)
    q = gen.dense_uniform_qubo(n, seed=12)
    sched = geo(500, 0.3, 6.0)
    res, _ = run_and_compare_dense(q, sched, 500, 21, capi.MODE_RANDOM_SITE, dtype,
                                   kernel=capi.KID_DENSE_GENERIC)
    assert res.stats["kernel_id"] == capi.KID_DENSE_GENERIC


def test_sharding_by_first_try_is_exact(gpu):
    """Global trajectory ids key the RNG: shards reproduce the unsharded run (multi-GPU row)."""
    q = gen.dense_uniform_qubo(96, seed=3)
    sched = geo(3, 0.3, 4.0)
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as prob:
        full = prob.anneal(sched, 3, 50, mode=capi.MODE_SEQUENTIAL_SWEEP, want_energies=True,
                           want_states=True)
        a = prob.anneal(sched, 3, 21, mode=capi.MODE_SEQUENTIAL_SWEEP, want_energies=True,
                        want_states=True)
        b = prob.anneal(sched, 3, 29, first_try=21, mode=capi.MODE_SEQUENTIAL_SWEEP,
                        want_energies=True, want_states=True)
        again = prob.anneal(sched, 3, 50, mode=capi.MODE_SEQUENTIAL_SWEEP, want_energies=True,
                            want_states=True)
    np.testing.assert_array_equal(np.vstack([a.best_states_packed, b.best_states_packed]),
                                  full.best_states_packed)
    np.testing.assert_array_equal(np.concatenate([a.best_energies, b.best_energies]),
                                  full.best_energies)
    np.testing.assert_array_equal(again.best_states_packed, full.best_states_packed)  # determinism
    k = min((a, b), key=lambda r: (r.energy, r.index))
    assert (k.energy, k.index) == (full.energy, full.index)


# ---------------------------------------------------------------- sparse CSR
@pytest.mark.parametrize("n,deg,dtype,mode,tries", [
    (128, 6, np.float64, capi.MODE_SEQUENTIAL_SWEEP, 70),
    (128, 6, np.float32, capi.MODE_RANDOM_SITE, 45),
    (517, 15, np.float32, capi.MODE_SEQUENTIAL_SWEEP, 100),
    (517, 15, np.float64, capi.MODE_RANDOM_SITE, 64),
    # sparse enough for groups of eight sites (osa_api.cu picks G = 8), ragged N, both precisions
    (3001, 7, np.float64, capi.MODE_SEQUENTIAL_SWEEP, 40),
    (3001, 7, np.float32, capi.MODE_SEQUENTIAL_SWEEP, 70),
    # degree 40: a half block no longer fits the staging buffer, the entries are read from global memory
    (300, 40, np.float32, capi.MODE_SEQUENTIAL_SWEEP, 50),
    (300, 40, np.float64, capi.MODE_SEQUENTIAL_SWEEP, 33),
])
def test_sparse_bit_exact(gpu, n, deg, dtype, mode, tries):
    rowptr, col, val, diag = gen.sparse_random_graph(n, deg, seed=n + deg)
    iters = 5 if mode == capi.MODE_SEQUENTIAL_SWEEP else 2000
    sched = geo(iters, 0.05, 1.5)
    prec = capi.SWEEP_F32 if dtype == np.float32 else capi.SWEEP_F64
    with Problem.csr(rowptr, col, val, diag, sweep_precision=prec) as prob:
        res = prob.anneal(sched, iters, tries, mode=mode, want_energies=True, want_states=True,
                          want_trace=True)
    with ob.trace(tries) as tr:
        best_rel, best, _, cnt = ob.replay_csr(rowptr, col, val, diag, sched, iters, tries,
                                               mode=mode, dtype=dtype)
    assert res.stats["kernel_id"] == capi.KID_SPARSE
    np.testing.assert_array_equal(res.trace_hash, tr.hashes)  # the spin sequence, flip by flip
    assert_states_equal(res.best_states_packed, best, "sparse best states vs host replay")
    assert res.stats["accepts"] == cnt.accepts
    q = gen.csr_to_dense(rowptr, col, val, diag)
    e_ref = ob.energy_packed(q, best)
    np.testing.assert_allclose(res.best_energies, e_ref, rtol=REL, atol=1e-12)
    k = int(np.argmin(e_ref))
    assert res.index == k and abs(res.energy - e_ref[k]) <= REL * max(1.0, abs(e_ref[k]))


def test_sparse_equals_dense_engine_on_integer_instance(gpu):
    """Same instance through the CSR and the dense kernels: with integer coefficients every
    partial sum is exact, so both layouts must walk identical trajectories."""
    rowptr, col, val, diag = gen.sparse_random_graph(96, 5, seed=9, integer=True)
    q = gen.csr_to_dense(rowptr, col, val, diag)
    sched = geo(4, 0.5, 8.0)
    with Problem.csr(rowptr, col, val, diag) as ps, Problem.dense(q) as pd:
        a = ps.anneal(sched, 4, 64, mode=capi.MODE_SEQUENTIAL_SWEEP, want_states=True)
        b = pd.anneal(sched, 4, 64, mode=capi.MODE_SEQUENTIAL_SWEEP, want_states=True)
    np.testing.assert_array_equal(a.best_states_packed, b.best_states_packed)
    assert a.energy == b.energy and a.index == b.index


# ---------------------------------------------------------------- energy + errors
def test_energy_batch_matches_reference_formula(gpu):
    rng = np.random.default_rng(0)
    for n in (1, 31, 32, 33, 600, 1025):
        q = gen.dense_uniform_qubo(n, seed=n)
        states = rng.integers(0, 2, size=(67, n)).astype(np.uint8)
        states[0] = 0
        states[1] = 1
        packed = pack_states(states)
        with Problem.dense(q) as prob:
            got = prob.energy_batch(packed)
        want = np.array([ob.ref_energy(q, s.astype(np.int8)) for s in states])
        np.testing.assert_allclose(got, want, rtol=REL, atol=1e-12)


def test_error_paths(gpu):
    q = gen.dense_uniform_qubo(8, seed=1)
    bad = q.copy()
    bad[0, 1] += 1.0
    with pytest.raises(capi.OsaError) as ei:
        Problem.dense(bad)
    assert ei.value.code == capi.OSA_ERR_INVALID
    with Problem.dense(q) as prob:
        with pytest.raises(capi.OsaError):
            prob.anneal(np.array([0.1, -1.0]), 2, 4)            # non-positive beta
        with pytest.raises(capi.OsaError):
            prob.anneal(geo(2), 2, 4, kernel_variant=capi.KID_SPARSE)
        with pytest.raises(capi.OsaError):
            prob.anneal(geo(2), 2, 4, kernel_variant=capi.KID_DENSE_SEQ)  # random mode


# ---------------------------------------------------------------- exhaustive search (next row 1)
def test_cuda_exhaustive_matches_reference_restatement(gpu):
    """osa_exhaustive_dense_f64 vs the oracle's restatement of exhaustive.hpp: same ground energy,
    same (lowest-index) ground state, on the reference's test instances and random ones."""
    import json
    import os
    from onesolver_b200 import exhaustive
    from oracle.qubo_format import parse_qubo
    here = os.path.dirname(os.path.abspath(__file__))
    g = json.load(open(os.path.join(here, "golden", "reference_vectors.json")))
    for case in g["exhaustive_test"]:
        n, lin, quad = parse_qubo(case["qubo"])
        q = ob.ref_flatten(n, lin, quad).reshape(n, n)
        state, e = exhaustive(q)
        want_state, want_e = ob.ref_exhaustive(q, n, 7)
        assert abs(e - case["energy"]) < 1e-13 and e == want_e
        np.testing.assert_array_equal(state, want_state)
    for n, seed, integer in [(1, 0, True), (2, 1, False), (17, 2, True), (19, 3, False),
                             (22, 4, True), (24, 2026, True), (25, 5, False)]:
        q = gen.dense_integer_qubo(n, seed) if integer else gen.dense_uniform_qubo(n, seed)
        state, e = exhaustive(q)
        want_state, want_e = ob.ref_exhaustive(q, n, 8)
        assert e == want_e, (n, e, want_e)
        np.testing.assert_array_equal(state, want_state)
    # degenerate instance: all states of the zero matrix tie -> lowest state integer (all zeros)
    state, e = exhaustive(np.zeros((12, 12)))
    assert e == 0.0 and not state.any()


def test_cli_gpu_paths(gpu, tmp_path):
    """one-solver-anneal / one-solver-exhaustive with --device-type gpu (BASELINE config 1 on GPU)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "app"), "-s", "-j4"], check=True)
    out = tmp_path / "a.csv"
    r = subprocess.run([os.path.join(root, "build/bin/one-solver-anneal"), "--input",
                        "examples/test1.qubo", "--output", str(out), "--device-type", "gpu"],
                       capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stderr
    assert "Using device: NVIDIA" in r.stdout
    assert out.read_text() == "0,1,2,3,energy\n1,1,0,1,-12\n"
    # host and gpu device types run the same engine: identical output on a Chimera instance
    args = ["--input", "tests/golden/chimera128/001.qubo", "--num-iter", "300", "--num-tries", "40",
            "--schedule-type", "linear", "--beta-max", "10"]
    outs = {}
    for dev in ("gpu", "cpu"):
        o = tmp_path / f"{dev}.csv"
        r = subprocess.run([os.path.join(root, "build/bin/one-solver-anneal")] + args +
                           ["--output", str(o), "--device-type", dev], capture_output=True,
                           text=True, cwd=root)
        assert r.returncode == 0, r.stderr
        outs[dev] = o.read_text()
    assert outs["gpu"] == outs["cpu"]
    ex = tmp_path / "e.csv"
    r = subprocess.run([os.path.join(root, "build/bin/one-solver-exhaustive"), "--input",
                        "examples/csp13.qubo", "--output", str(ex), "--device-type", "gpu"],
                       capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stderr
    assert ex.read_text() == "0,1,2,3,4,5,6,7,8,9,10,11,12,energy\n1,1,1,0,1,1,0,0,0,0,0,0,0,-32\n"


def test_cli_csr_route(gpu, tmp_path):
    """The shim's CSR route (include/simulated_annealing/annealing.hpp: Layout::automatic picks
    CSR for N > 2048 and density < 0.05; --layout csr forces it) through one-solver-anneal on the
    GPU: the sparse kernel runs, and the CSV equals the dense route's and the host engine's."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-C", os.path.join(root, "app"), "-s", "-j4"], check=True)
    exe = os.path.join(root, "build/bin/one-solver-anneal")
    # generated sparse instance, N = 2600 (> 2048), 7800 couplers (density 0.0023), integer
    # coefficients: every partial sum is exact, so all layouts and engines walk the same path
    big = tmp_path / "sparse2600.qubo"
    r = subprocess.run([os.path.join(root, "build/bin/qubo-io-bench"), "--generate", str(big),
                        "2600", "7800", "11"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

    def cli(inp, dev, *extra):
        o = tmp_path / f"out_{dev}_{len(extra)}_{abs(hash(extra)) % 9973}.csv"
        r = subprocess.run([exe, "--input", str(inp), "--output", str(o), "--device-type", dev,
                            "--num-iter", "12", "--num-tries", "48", "--mode", "sweep",
                            "--schedule-type", "geometric", "--beta-min", "0.05", "--beta-max", "3",
                            "--stats"] + list(extra), capture_output=True, text=True, cwd=root)
        assert r.returncode == 0, r.stderr + r.stdout
        return o.read_text(), r.stdout

    auto_csv, auto_out = cli(big, "gpu", "--num-gpus", "1")
    assert "Kernel: sparse_csr" in auto_out          # Layout::automatic -> CSR
    dense_csv, dense_out = cli(big, "gpu", "--num-gpus", "1", "--layout", "dense")
    assert "Kernel: dense_seq" in dense_out
    host_csv, _ = cli(big, "cpu")
    assert auto_csv == dense_csv == host_csv
    assert auto_csv.count("\n") == 2 and auto_csv.splitlines()[0].endswith(",energy")
    # fp32 fields on the CSR route: integer instance -> still the same walk
    f32_csv, f32_out = cli(big, "gpu", "--num-gpus", "1", "--precision", "f32")
    assert "Kernel: sparse_csr" in f32_out and f32_csv == auto_csv
    # Chimera-512 (N = 512: automatic stays dense) forced onto the CSR kernel
    chim = os.path.join(root, "tests/golden/chimera512/001.qubo")
    csr_csv, csr_out = cli(chim, "gpu", "--num-gpus", "1", "--layout", "csr")
    assert "Kernel: sparse_csr" in csr_out
    ref_csv, ref_out = cli(chim, "gpu", "--num-gpus", "1")
    assert "Kernel: dense_seq" in ref_out
    host_chim, _ = cli(chim, "cpu")
    assert ref_csv == host_chim
    # fractional couplings: the CSR kernel sums a field in CSR order, the dense engines update it
    # flip by flip, so the two may differ in the last bits of a field; the energies of the
    # returned states must agree to the printed precision or be a different local optimum of
    # comparable quality (same schedule, same streams)
    e_csr = float(csr_csv.splitlines()[1].rsplit(",", 1)[1])
    e_ref = float(ref_csv.splitlines()[1].rsplit(",", 1)[1])
    assert abs(e_csr - e_ref) <= 0.05 * abs(e_ref)


@pytest.mark.parametrize("n,dtype,tries", [(200, np.float32, 30), (1100, np.float32, 17),
                                           (513, np.float64, 20), (4096, np.float32, 13),
                                           (2048, np.float64, 9), (5000, np.float32, 9)])
def test_single_role_kernel_is_bit_identical_to_warp_specialised(gpu, monkeypatch, n, dtype, tries):
    """OSA_DS_WS=0 selects k_dense_seq (decide and apply back to back); the default k_dense_seq_ws
    overlaps them.  Both must walk exactly the trajectories of the host replay."""
    q = gen.dense_uniform_qubo(n, seed=300 + n)
    scale = np.sqrt(n)
    sched = geo(3, 0.02 * scale, 0.6 * scale)
    a, _ = run_and_compare_dense(q, sched, 3, tries, capi.MODE_SEQUENTIAL_SWEEP, dtype)
    # the other warp-specialised variant (lock-step roles <-> free-running roles): the default
    # picks the free-running kernel for fp32 rows of >= 4096 and fp64 rows of >= 1536 elements
    # (osa_dense_seq.cu, use_flow)
    unit = 1024 if dtype == np.float32 else 512
    ld = -(-n // unit) * unit
    default_is_flow = ld >= (4096 if dtype == np.float32 else 1536)
    monkeypatch.setenv("OSA_WS_FLOW", "0" if default_is_flow else "1")
    c, _ = run_and_compare_dense(q, sched, 3, tries, capi.MODE_SEQUENTIAL_SWEEP, dtype)
    monkeypatch.delenv("OSA_WS_FLOW")
    monkeypatch.setenv("OSA_DS_WS", "0")
    b, _ = run_and_compare_dense(q, sched, 3, tries, capi.MODE_SEQUENTIAL_SWEEP, dtype)
    np.testing.assert_array_equal(a.best_states_packed, b.best_states_packed)
    np.testing.assert_array_equal(a.best_states_packed, c.best_states_packed)
    np.testing.assert_array_equal(a.trace_hash, b.trace_hash)
    np.testing.assert_array_equal(a.trace_hash, c.trace_hash)
    assert a.stats["grid"] != 0 and b.stats["grid"] != 0


def test_csr_energy_kernels_agree_bit_for_bit(gpu, monkeypatch):
    """Exact fp64 energies on a CSR problem: 32 states per warp (staged transposed in shared memory,
    a site's CSR row read once per warp) against one state per thread (OSA_ENERGY_CSR_SCALAR=1) --
    the same additions in the same order, so the same bits, and both equal to the reference formula
    to rounding."""
    for n, deg in ((1500, 11), (77, 5), (3001, 3)):
        rowptr, col, val, diag = gen.sparse_random_graph(n, deg, seed=77 + n)
        rng = np.random.default_rng(3)
        states = pack_states(rng.integers(0, 2, size=(333, n)).astype(np.uint8))
        with Problem.csr(rowptr, col, val, diag) as prob:
            monkeypatch.delenv("OSA_ENERGY_CSR_SCALAR", raising=False)
            a = prob.energy_batch(states)
            monkeypatch.setenv("OSA_ENERGY_CSR_SCALAR", "1")
            b = prob.energy_batch(states)
        monkeypatch.delenv("OSA_ENERGY_CSR_SCALAR", raising=False)
        np.testing.assert_array_equal(a, b)
        q = gen.csr_to_dense(rowptr, col, val, diag)
        np.testing.assert_allclose(a, ob.energy_packed(q, states), rtol=REL, atol=1e-9)
