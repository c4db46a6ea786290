"""One dense fp64 N=1024 run (config 3 shape, SWEEPS sweeps of its schedule's cold end) for profiling."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from onesolver_b200 import Problem, capi
from onesolver_b200 import problems as gen
n, tries, sweeps = 1024, int(os.environ.get("TRIES", 16384)), int(os.environ.get("SWEEPS", 20))
q = gen.dense_uniform_qubo(n, seed=2024 + 3)
sched = 0.64 * (9.6 / 0.64) ** (np.arange(sweeps) / max(1, sweeps - 1))
with Problem.dense(q, sweep_precision=capi.SWEEP_F64) as p:
    r = p.anneal(sched, sweeps, tries, mode=capi.MODE_SEQUENTIAL_SWEEP)
    print(r.stats["kernel_id"], r.stats["ms_sweep"], r.stats["accepts"] / r.stats["attempts"], r.energy)
