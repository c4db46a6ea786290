"""Time the exact-energy kernel (osa_energy_batch) on random states: N=4096, 131072 states."""
import sys, time, os
sys.path.insert(0, ".")
import numpy as np
from onesolver_b200 import Problem, capi
from onesolver_b200 import problems as gen
n = int(os.environ.get("N", 4096)); count = int(os.environ.get("COUNT", 131072))
q = gen.dense_uniform_qubo(n, seed=3)
rng = np.random.default_rng(1)
states = rng.integers(0, 2**32, size=(count, (n + 31) // 32), dtype=np.uint32)
with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as p:
    for rep in range(3):
        t0 = time.perf_counter()
        e = p.energy_batch(states)
        t1 = time.perf_counter()
        print("energy_batch %d states N=%d: %.1f ms wall (incl. %.0f MB H2D)" % (count, n, (t1 - t0) * 1e3, states.nbytes / 1e6), flush=True)
    # spot check against numpy on a few states
    from onesolver_b200 import unpack_states
    x = unpack_states(states[:4], n).astype(np.float64)
    ref = [float(xi @ np.triu(q) @ xi) for xi in x]
    print("check", np.max(np.abs(np.array(ref) - e[:4]) / np.abs(ref)))
