"""Seeded synthetic instances shared by the tests, smoke() and bench.py."""
import numpy as np


def dense_integer_qubo(n, seed, lo=-10, hi=10):
    """flatten_qubo layout, integer-valued (exactly representable) coefficients."""
    rng = np.random.default_rng(seed)
    a = rng.integers(lo, hi + 1, size=(n, n)).astype(np.float64)
    q = np.triu(a, 1)
    q = q + q.T
    q[np.arange(n), np.arange(n)] = rng.integers(lo, hi + 1, size=n)
    return q


def dense_uniform_qubo(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1.0, 1.0, size=(n, n))
    q = np.triu(a, 1)
    q = q + q.T
    q[np.arange(n), np.arange(n)] = rng.uniform(-1.0, 1.0, size=n)
    return q


def sparse_random_graph(n, degree, seed, integer=False):
    """Symmetric random graph with max degree <= `degree`; returns CSR + diag + dense copy."""
    rng = np.random.default_rng(seed)
    nbrs = [dict() for _ in range(n)]
    target_edges = n * degree // 2
    tries = 0
    edges = 0
    while edges < target_edges and tries < 20 * target_edges:
        tries += 1
        i, j = int(rng.integers(0, n)), int(rng.integers(0, n))
        if i == j or j in nbrs[i] or len(nbrs[i]) >= degree or len(nbrs[j]) >= degree:
            continue
        v = float(rng.integers(-5, 6)) if integer else float(rng.uniform(-1, 1))
        nbrs[i][j] = v
        nbrs[j][i] = v
        edges += 1
    diag = (rng.integers(-5, 6, size=n).astype(np.float64) if integer
            else rng.uniform(-1, 1, size=n))
    rowptr = np.zeros(n + 1, dtype=np.int32)
    col, val = [], []
    for i in range(n):
        for j in sorted(nbrs[i]):
            col.append(j)
            val.append(nbrs[i][j])
        rowptr[i + 1] = len(col)
    col = np.array(col, dtype=np.int32)
    val = np.array(val, dtype=np.float64)
    return rowptr, col, val, diag


def csr_to_dense(rowptr, col, val, diag):
    n = len(diag)
    q = np.zeros((n, n), dtype=np.float64)
    for i in range(n):
        for p in range(rowptr[i], rowptr[i + 1]):
            q[i, col[p]] = val[p]
    q[np.arange(n), np.arange(n)] = diag
    return q
