"""GPU tests at BASELINE.json's full shapes (configs 3, 4, 5): too large for a full oracle run, so
they check size-independent properties plus an oracle replay of a sample of the trajectories.

Properties: returned (state, energy) self-consistent under the reference energy function, winner =
first minimum of the per-trajectory energies, prefix stability in num_tries, shard
invariance, counters consistent, and sampled trajectories bit-exact vs the host replay."""
import numpy as np
import pytest

from onesolver_b200 import Problem, capi, unpack_states
from onesolver_b200 import problems as gen
from oracle import binding as ob

pytestmark = pytest.mark.gpu
REL = 1e-9


def check_common(res, tries):
    e = res.best_energies
    k = int(np.argmin(e))
    assert res.index == k and res.energy == e[k]
    assert res.stats["attempts"] > 0 and 0 < res.stats["accepts"] < res.stats["attempts"]
    assert res.stats["row_fetches"] <= res.stats["accepts"]
    assert e.shape == (tries,) and np.isfinite(e).all()


def test_config3_dense_fp64_n1024_16384_tries(gpu):
    """BASELINE config 3 shape: dense fp64 N=1024, 16384 tries (12 sweeps here; the sweep count
    only scales run time)."""
    n, tries, sweeps = 1024, 16384, 12
    q = gen.dense_uniform_qubo(n, seed=2024 + 3)
    sched = ob.ref_schedule("geometric", 0.6, 10.0, sweeps)
    with Problem.dense(q, sweep_precision=capi.SWEEP_F64) as prob:
        res = prob.anneal(sched, sweeps, tries, mode=capi.MODE_SEQUENTIAL_SWEEP,
                          want_energies=True, want_states=True)
        small = prob.anneal(sched, sweeps, 300, mode=capi.MODE_SEQUENTIAL_SWEEP,
                            want_energies=True, want_states=True)
        tail = prob.anneal(sched, sweeps, 100, first_try=tries - 100,
                           mode=capi.MODE_SEQUENTIAL_SWEEP, want_states=True)
    check_common(res, tries)
    assert res.stats["kernel_id"] == capi.KID_DENSE_SEQ and res.stats["q_elem_bytes"] == 8
    # prefix stability and shard invariance
    np.testing.assert_array_equal(small.best_states_packed, res.best_states_packed[:300])
    np.testing.assert_array_equal(small.best_energies, res.best_energies[:300])
    np.testing.assert_array_equal(tail.best_states_packed, res.best_states_packed[-100:])
    # sampled trajectories vs the oracle: replay ids 0..23 and 16360..16383
    for first in (0, tries - 24):
        _, best, _, _ = ob.replay_dense(q, sched, sweeps, 24, mode=1, first_try=first,
                                        dtype=np.float64)
        np.testing.assert_array_equal(res.best_states_packed[first:first + 24], best)
        e_ref = ob.energy_packed(q, best)
        np.testing.assert_allclose(res.best_energies[first:first + 24], e_ref, rtol=REL)
    # the winner under the reference energy function
    e_win = ob.ref_energy(q, res.state.astype(np.int8))
    assert abs(e_win - res.energy) <= REL * abs(e_win)


def test_config5_dense_n4096_131072_tries_per_gpu(gpu):
    """BASELINE config 5 per-GPU share: dense N=4096, 131072 tries (1 sweep here)."""
    n, tries = 4096, 131072
    q = gen.dense_uniform_qubo(n, seed=2024 + 5)
    sched = np.array([8.0])
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as prob:
        res = prob.anneal(sched, 1, tries, mode=capi.MODE_SEQUENTIAL_SWEEP, want_energies=True,
                          want_states=True)
        shard = prob.anneal(sched, 1, 64, first_try=70000, mode=capi.MODE_SEQUENTIAL_SWEEP,
                            want_energies=True, want_states=True)
    check_common(res, tries)
    assert res.stats["traj_per_batch"] >= 8 and res.stats["q_elem_bytes"] == 4
    assert res.stats["attempts"] == tries * n
    np.testing.assert_array_equal(shard.best_states_packed, res.best_states_packed[70000:70064])
    np.testing.assert_array_equal(shard.best_energies, res.best_energies[70000:70064])
    r = res.stats["traj_per_batch"]
    first = 131072 - 2 * r
    _, best, _, cnt = ob.replay_dense(q, sched, 1, 2 * r, mode=1, first_try=first, dtype=np.float32,
                                      batch_r=r)
    np.testing.assert_array_equal(res.best_states_packed[first:], best)
    e_ref = ob.energy_packed(q, best)
    np.testing.assert_allclose(res.best_energies[first:], e_ref, rtol=REL)
    e_win = ob.ref_energy(q, res.state.astype(np.int8))
    assert abs(e_win - res.energy) <= REL * abs(e_win)
    assert (unpack_states(res.best_states_packed[res.index], n)[0] == res.state).all()


def test_config5_bench_schedule_replayed_from_the_middle_of_the_id_range(gpu):
    """The bench workload itself (bench.py defaults: 32 sweeps, geometric beta 1.28 -> 19.2,
    131072 tries, fp32 fields), and two whole CTAs' worth of trajectories from the MIDDLE of the id
    range replayed on the host: flip traces (the complete spin sequence), best states, energies."""
    import argparse
    import bench
    a = argparse.Namespace(n=4096, sweeps=32, beta_min=1.28, beta_max=19.2)
    n, tries = 4096, 131072
    q = bench.make_instance(n)
    sched = bench.make_schedule(a)
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as prob:
        res = prob.anneal(sched, 32, tries, mode=capi.MODE_SEQUENTIAL_SWEEP, want_energies=True,
                          want_states=True, want_trace=True)
    check_common(res, tries)
    assert res.stats["attempts"] == tries * n * 32
    assert abs(res.stats["accepts"] / res.stats["attempts"] - 0.171) < 0.005  # the frozen workload
    r = res.stats["traj_per_batch"]
    first = (tries // 2 // r) * r  # a batch boundary in the middle
    with ob.trace(2 * r) as tr:
        _, best, _, _ = ob.replay_dense(q, sched, 32, 2 * r, mode=1, first_try=first,
                                        dtype=np.float32, batch_r=r)
    np.testing.assert_array_equal(res.trace_hash[first:first + 2 * r], tr.hashes)
    np.testing.assert_array_equal(res.best_states_packed[first:first + 2 * r], best)
    e_ref = ob.energy_packed(q, best)
    np.testing.assert_allclose(res.best_energies[first:first + 2 * r], e_ref, rtol=REL)
    # fp32 sweep, energies re-scored in fp64: north_star's 1e-5 bar for fp32 is met with room
    e_win = ob.ref_energy(q, res.state.astype(np.int8))
    assert abs(e_win - res.energy) <= REL * abs(e_win)


def test_config4_sparse_n5627_65536_tries_linear_schedule(gpu):
    """BASELINE config 4 shape: sparse degree<=15 QUBO on N=5627 (CSR), 65536 tries, linear schedule."""
    n, tries, sweeps = 5627, 65536, 3
    rowptr, col, val, diag = gen.sparse_random_graph(n, 15, seed=2024 + 4)
    sched = ob.ref_schedule("linear", 0.05, 1.0, sweeps)
    with Problem.csr(rowptr, col, val, diag, sweep_precision=capi.SWEEP_F32) as prob:
        res = prob.anneal(sched, sweeps, tries, mode=capi.MODE_SEQUENTIAL_SWEEP,
                          want_energies=True, want_states=True)
        shard = prob.anneal(sched, sweeps, 96, first_try=4000, mode=capi.MODE_SEQUENTIAL_SWEEP,
                            want_states=True)
    e = res.best_energies
    k = int(np.argmin(e))
    assert res.index == k and res.energy == e[k]
    assert res.stats["kernel_id"] == capi.KID_SPARSE
    np.testing.assert_array_equal(shard.best_states_packed, res.best_states_packed[4000:4096])
    for first in (0, tries - 40):
        _, best, _, _ = ob.replay_csr(rowptr, col, val, diag, sched, sweeps, 40, mode=1,
                                      first_try=first, dtype=np.float32)
        np.testing.assert_array_equal(res.best_states_packed[first:first + 40], best)
        # reference energy function on the sparse instance: diag.x + sum_{i<j} q_ij x_i x_j
        x = unpack_states(best, n).astype(np.float64)
        rows = np.repeat(np.arange(n), np.diff(rowptr))
        upper = col > rows
        e_ref = x @ diag + ((x[:, rows[upper]] * x[:, col[upper]]) @ val[upper])
        np.testing.assert_allclose(res.best_energies[first:first + 40], e_ref, rtol=REL, atol=1e-9)


def test_dense_beyond_the_register_kernel_n10000(gpu):
    """N = 10000 > 8192: sequential sweeps fall through to the warp-per-trajectory kernel (fields in
    shared memory); bit-exact against the host replay, energies against the reference formula."""
    n, tries = 10000, 6
    q = gen.dense_uniform_qubo(n, seed=77)
    sched = np.array([12.0, 30.0])
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as prob:
        res = prob.anneal(sched, 2, tries, mode=capi.MODE_SEQUENTIAL_SWEEP, want_energies=True,
                          want_states=True, want_trace=True)
    assert res.stats["kernel_id"] == capi.KID_DENSE_GENERIC
    with ob.trace(tries) as tr:
        _, best, _, cnt = ob.replay_dense(q, sched, 2, tries, mode=1, dtype=np.float32, batch_r=1)
    np.testing.assert_array_equal(res.trace_hash, tr.hashes)
    np.testing.assert_array_equal(res.best_states_packed, best)
    assert res.stats["accepts"] == cnt.accepts
    np.testing.assert_allclose(res.best_energies, ob.energy_packed(q, best), rtol=REL)


@pytest.mark.parametrize("n,dtype", [(57000, np.float32), (56000, np.float64)])
def test_sparse_at_the_largest_supported_size(gpu, n, dtype):
    """The sparse kernel keeps the spin words of 32 trajectories in shared memory: N up to ~57k.
    One warp per SM at this size; bit-exact against the host replay, and one site more is refused."""
    rowptr, col, val, diag = gen.sparse_random_graph(n, 3, seed=n)
    sched = ob.ref_schedule("linear", 0.1, 2.0, 2)
    prec = capi.SWEEP_F32 if dtype == np.float32 else capi.SWEEP_F64
    with Problem.csr(rowptr, col, val, diag, sweep_precision=prec) as prob:
        res = prob.anneal(sched, 2, 40, mode=capi.MODE_SEQUENTIAL_SWEEP, want_states=True,
                          want_trace=True)
    with ob.trace(40) as tr:
        _, best, _, cnt = ob.replay_csr(rowptr, col, val, diag, sched, 2, 40, mode=1, dtype=dtype)
    np.testing.assert_array_equal(res.trace_hash, tr.hashes)
    np.testing.assert_array_equal(res.best_states_packed, best)
    assert res.stats["accepts"] == cnt.accepts
    big = 57100 if dtype == np.float32 else 56100
    rp = np.zeros(big + 1, dtype=np.int32)
    with pytest.raises(capi.OsaError, match="sparse kernel supports"):
        Problem.csr(rp, np.zeros(0, dtype=np.int32), np.zeros(0), np.zeros(big), sweep_precision=prec)
