"""osa_multi_anneal on G devices from one plain process (no torch): per-device kernel times and the
wall clock of the call, for the bench-shaped workload on 4 waves of trajectories per GPU.
Usage: python tools/multi_probe.py [G ...]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from onesolver_b200 import MultiProblem, capi, device_count  # noqa: E402
from onesolver_b200 import problems as gen  # noqa: E402

q = gen.dense_uniform_qubo(4096, seed=2029)
sched = np.geomspace(1.28, 19.2, 32)
for g in [int(a) for a in sys.argv[1:]] or [1, device_count()]:
    tries = 148 * 12 * 4 * g
    with MultiProblem.dense(q, num_devices=g, sweep_precision=capi.SWEEP_F32) as m:
        for rep in range(3):
            t0 = time.perf_counter()
            r = m.anneal(sched, 32, tries, mode=capi.MODE_SEQUENTIAL_SWEEP)
            wall = (time.perf_counter() - t0) * 1e3
        print(f"G={g} tries={tries} wall {wall:.1f} ms, per-device sweep ms",
              [round(d["ms_sweep"], 1) for d in r.device_stats], "attempts/s %.3e" % (r.stats["attempts"] / wall * 1e3),
              "best", r.energy, r.index)
