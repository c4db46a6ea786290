"""One CTA of the N=4096 fp32 instantiation of the sweep kernel (the bench shape) for compute-sanitizer."""
import sys
import numpy as np
sys.path.insert(0, ".")
from onesolver_b200 import Problem, capi
from onesolver_b200 import problems as gen
q = gen.dense_uniform_qubo(4096, seed=7)
with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as p:
    r = p.anneal(np.array([6.0]), 1, 12, mode=capi.MODE_SEQUENTIAL_SWEEP)
print("ok", r.stats["kernel_id"], r.stats["traj_per_batch"], r.stats["accepts"], r.energy)
