"""Seeded fuzz of the sparse sweep kernel's grouped layout (osa_sparse.cu, built in osa_api.cu) against
the host replay: sizes around the group / block / staging boundaries, empty rows, hub sites whose
degree exceeds the staging buffer, dense-ish graphs (groups of four) and very sparse ones (groups of
eight), both precisions, both modes.  Bit-exact best states, flip traces and accept counters."""
import numpy as np
import pytest

from onesolver_b200 import Problem, capi
from onesolver_b200 import problems as gen
from oracle import binding as ob

pytestmark = pytest.mark.gpu


def random_symmetric_csr(n, edges, seed, hubs=0, integer=False):
    """Random simple graph with `edges` couplers plus `hubs` sites joined to half of all sites."""
    rng = np.random.default_rng(seed)
    nb = [dict() for _ in range(n)]

    def add(i, j):
        if i == j or j in nb[i]:
            return
        v = float(rng.integers(-4, 5)) if integer else float(rng.uniform(-1, 1))
        nb[i][j] = v
        nb[j][i] = v

    for _ in range(edges):
        add(int(rng.integers(0, n)), int(rng.integers(0, n)))
    for h in range(hubs):
        hub = int(rng.integers(0, n))
        for j in rng.choice(n, size=max(1, n // 2), replace=False):
            add(hub, int(j))
    rowptr = np.zeros(n + 1, dtype=np.int32)
    col, val = [], []
    for i in range(n):
        for j in sorted(nb[i]):
            col.append(j)
            val.append(nb[i][j])
        rowptr[i + 1] = len(col)
    diag = rng.integers(-4, 5, size=n).astype(np.float64) if integer else rng.uniform(-1, 1, size=n)
    return rowptr, np.array(col, dtype=np.int32), np.array(val, dtype=np.float64), diag


CASES = []
_rng = np.random.default_rng(20261018)
for _n in (1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 95, 129, 255, 257, 800, 1201, 2050):
    CASES.append((_n, int(_rng.integers(0, 4 * _n + 1)), 0))
CASES += [(400, 200, 1), (1500, 700, 2), (900, 40000, 0), (64, 2016, 0), (3000, 0, 0), (777, 6000, 3)]


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_sparse_layout_fuzz(gpu, idx):
    n, edges, hubs = CASES[idx]
    dtype = np.float32 if idx % 2 else np.float64
    rowptr, col, val, diag = random_symmetric_csr(n, edges, seed=1000 + idx, hubs=hubs,
                                                  integer=(idx % 5 == 0))
    sweeps, tries = 3, 33 + (idx % 3) * 31
    sched = ob.ref_schedule("geometric", 0.05, 2.0, sweeps)
    prec = capi.SWEEP_F32 if dtype == np.float32 else capi.SWEEP_F64
    with Problem.csr(rowptr, col, val, diag, sweep_precision=prec) as prob:
        res = prob.anneal(sched, sweeps, tries, mode=capi.MODE_SEQUENTIAL_SWEEP, want_states=True,
                          want_trace=True, first_try=idx)
        rnd = prob.anneal(np.full(50, 0.7), 50, tries, mode=capi.MODE_RANDOM_SITE, want_states=True,
                          want_trace=True)
    with ob.trace(tries) as tr:
        _, best, _, cnt = ob.replay_csr(rowptr, col, val, diag, sched, sweeps, tries, mode=1,
                                        dtype=dtype, first_try=idx)
    np.testing.assert_array_equal(res.trace_hash, tr.hashes)
    np.testing.assert_array_equal(res.best_states_packed, best)
    assert res.stats["accepts"] == cnt.accepts and res.stats["kernel_id"] == capi.KID_SPARSE
    with ob.trace(tries) as tr:
        _, best, _, cnt = ob.replay_csr(rowptr, col, val, diag, np.full(50, 0.7), 50, tries, mode=0,
                                        dtype=dtype)
    np.testing.assert_array_equal(rnd.trace_hash, tr.hashes)
    np.testing.assert_array_equal(rnd.best_states_packed, best)
