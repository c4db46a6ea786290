"""One random-site (reference loop) run on the headline instance for profiling: N=4096 fp32."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from onesolver_b200 import Problem, capi
from onesolver_b200 import problems as gen
n, tries, iters = 4096, int(os.environ.get("TRIES", 4736)), int(os.environ.get("ITERS", 1024))
q = gen.dense_uniform_qubo(n, seed=2024 + 5)
sched = 1.28 * (19.2 / 1.28) ** (np.arange(iters) / (iters - 1))
with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as p:
    r = p.anneal(sched, iters, tries, mode=capi.MODE_RANDOM_SITE)
    print(r.stats["kernel_id"], r.stats["ms_sweep"], r.energy)
