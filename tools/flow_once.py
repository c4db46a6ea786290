"""One wave of the headline workload (N=4096 fp32, bench schedule, 148 x 12 trajectories) for an ncu capture;
prints the row counters of the launch so that instructions per streamed row can be derived."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
from onesolver_b200 import Problem, capi
from onesolver_b200 import problems as gen
q = gen.dense_uniform_qubo(4096, seed=2024 + 5)
sched = 1.28 * (19.2 / 1.28) ** (np.arange(32) / 31.0)
with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as p:
    r = p.anneal(sched, 32, 148 * 12, mode=capi.MODE_SEQUENTIAL_SWEEP)
st = r.stats
print(json.dumps({"rows": st["row_fetches"] + st["init_row_fetches"], "row_fetches": st["row_fetches"],
                  "init_row_fetches": st["init_row_fetches"], "accepts": st["accepts"],
                  "attempts": st["attempts"], "ms_sweep": st["ms_sweep"], "grid": st["grid"]}))
