"""Phases of the end-to-end step of the random-site sub-record (create / anneal / destroy), as
bench.py runs it: a resident problem annealed W + K times first, then create + anneal + destroy."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from onesolver_b200 import MultiProblem, capi, construct_geometric_beta_schedule
from onesolver_b200 import problems as gen
n, tries, iters = 4096, 16384, 4096
q = gen.dense_uniform_qubo(n, seed=2024 + 5)
sched = construct_geometric_beta_schedule(1.28, 19.2, iters)
mk = lambda: MultiProblem.dense(q, devices=[0], sweep_precision=capi.SWEEP_F32)
with mk() as p:
    for i in range(4):
        t0 = time.perf_counter()
        r = p.anneal(sched, iters, tries, mode=capi.MODE_RANDOM_SITE)
        print("resident anneal %.1f ms wall, device %.1f" % ((time.perf_counter() - t0) * 1e3, r.stats["ms_total"]), flush=True)
for i in range(6):
    t0 = time.perf_counter()
    p = mk()
    t1 = time.perf_counter()
    r = p.anneal(sched, iters, tries, mode=capi.MODE_RANDOM_SITE, want_energies=True)
    t2 = time.perf_counter()
    p.close()
    t3 = time.perf_counter()
    print("e2e step: create %.1f ms anneal %.1f ms (device %.1f) destroy %.1f ms" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, r.stats["ms_total"], (t3 - t2) * 1e3), flush=True)
