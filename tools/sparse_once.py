"""One sparse (config 4 shape) run for profiling: N=5627, degree<=15, 65536 tries, SWEEPS sweeps."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from onesolver_b200 import Problem, capi
from onesolver_b200 import problems as gen
n = 5627
sweeps = int(os.environ.get("SWEEPS", 2))
rowptr, col, val, diag = gen.sparse_random_graph(n, 15, seed=2028)
with Problem.csr(rowptr, col, val, diag, sweep_precision=capi.SWEEP_F32) as p:
    r = p.anneal(np.linspace(0.5, 4.0, sweeps), sweeps, 65536, mode=capi.MODE_SEQUENTIAL_SWEEP)
    print(r.stats["ms_sweep"], r.energy)
