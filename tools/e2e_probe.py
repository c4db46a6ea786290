"""Where does an end-to-end sa::anneal call spend its time?  create / anneal (device ms) / destroy,
with the QUBO in pageable or pinned host memory, at a small and at the bench's trajectory count."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from onesolver_b200 import Problem, capi, pinned_copy
from onesolver_b200 import problems as gen
q = gen.dense_uniform_qubo(4096, seed=1)
qp, free = pinned_copy(q)
for tries, sweeps in ((1776, 2), (131072, 32)):
    sched = np.linspace(1.0, 2.0, sweeps)
    for name, arr in (("pageable", q), ("pinned", qp), ("pageable", q), ("pinned", qp)):
        t0 = time.perf_counter()
        p = Problem.dense(arr, sweep_precision=capi.SWEEP_F32)
        t1 = time.perf_counter()
        r = p.anneal(sched, sweeps, tries, mode=1, want_energies=True)
        t2 = time.perf_counter()
        p.close()
        t3 = time.perf_counter()
        print("%6d tries %-8s create %.1f ms anneal %.1f ms (device %.1f) destroy %.1f ms" % (
            tries, name, (t1-t0)*1e3, (t2-t1)*1e3, r.stats["ms_total"], (t3-t2)*1e3), flush=True)
free()
