"""GPU probe: read-bandwidth ladder (L2 vs HBM) and quick throughput of the sweep kernels.
Usage: python tools/probe.py [bw] [dense] [sparse]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from onesolver_b200 import Problem, capi, device_name, measure_read_bandwidth  # noqa: E402
from onesolver_b200 import problems as gen  # noqa: E402

what = set(sys.argv[1:]) or {"bw", "dense", "sparse"}
print("device:", device_name(0))

if "bw" in what:
    for mb in (8, 16, 32, 48, 64, 80, 96, 128, 256, 1024, 4096):
        iters = max(2, min(200, 8192 // mb))
        g = measure_read_bandwidth(mb << 20, iters)
        print(json.dumps({"probe": "read_bw", "mb": mb, "iters": iters, "gbs": round(g, 1)}))


def geo(n, lo, hi):
    a = (hi / lo) ** (1.0 / max(1, n - 1))
    return np.array([lo * a ** i for i in range(n)])


def run_dense(n, prec, tries, sweeps, lo, hi, label):
    q = gen.dense_uniform_qubo(n, seed=2024)
    t0 = time.time()
    with Problem.dense(q, sweep_precision=prec) as prob:
        t1 = time.time()
        sched = geo(sweeps, lo, hi)
        for rep in range(2):
            res = prob.anneal(sched, sweeps, tries, mode=capi.MODE_SEQUENTIAL_SWEEP)
        st = res.stats
    esz = st["q_elem_bytes"]
    ld = -(-n // (1024 if esz == 4 else 512)) * (1024 if esz == 4 else 512)
    sec = st["ms_sweep"] * 1e-3
    out = {"probe": label, "n": n, "tries": tries, "sweeps": sweeps, "R": st["traj_per_batch"],
           "grid": st["grid"], "ms_sweep": round(st["ms_sweep"], 3),
           "ms_energy": round(st["ms_energy"], 3), "ms_total": round(st["ms_total"], 3),
           "attempts_per_s": st["attempts"] / sec, "accept_frac": st["accepts"] / st["attempts"],
           "row_fetches": st["row_fetches"], "init_rows": st["init_row_fetches"],
           "row_gbs": (st["row_fetches"] + st["init_row_fetches"]) * ld * esz / sec / 1e9,
           "unshared_gbs": st["accepts"] * ld * esz / sec / 1e9,
           "upload_s": round(t1 - t0, 3), "energy": res.energy,
           "kcyc_per_cta": {k: round(st[k] / max(1, st["grid"]) / 1e3, 1)
                            for k in ("cyc_init", "cyc_stage", "cyc_decide", "cyc_apply")}}
    print(json.dumps(out))


if "dense" in what:
    s = np.sqrt(4096)
    run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 0.3 * s, 0.02 * s, "dense4096_f32_1wave_hot2cold")
    run_dense(4096, capi.SWEEP_F32, 148 * 12 * 4, 4, 0.3 * s, 0.02 * s, "dense4096_f32_4waves")
    run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 0.01 * s, 0.002 * s, "dense4096_f32_cold")
    run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 2 * s, 1 * s, "dense4096_f32_hot")
    s = np.sqrt(1024)
    run_dense(1024, capi.SWEEP_F64, 148 * 16, 8, 0.3 * s, 0.02 * s, "dense1024_f64")
    run_dense(1024, capi.SWEEP_F32, 148 * 16, 8, 0.3 * s, 0.02 * s, "dense1024_f32")
    s = np.sqrt(2048)
    run_dense(2048, capi.SWEEP_F32, 148 * 16, 4, 0.3 * s, 0.02 * s, "dense2048_f32")
    s = np.sqrt(512)
    run_dense(512, capi.SWEEP_F64, 148 * 16, 16, 0.3 * s, 0.02 * s, "dense512_f64")
    run_dense(4096, capi.SWEEP_F64, 148 * 4, 2, 0.3 * 64, 0.02 * 64, "dense4096_f64")

if "dbg" in what:
    # role experiments of the warp-specialised kernel (DenseParams::debug_flags): 8 = normal run,
    # cyc_init = walk iterations of decide warp 0; 9 = decide warps alone; 2 = apply warps alone
    import os
    s = np.sqrt(4096)
    for flags in ("8", "9", "2"):
        os.environ["OSA_WS_DEBUG"] = flags
        run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 0.3 * s, 0.02 * s, "dbg%s_4096_hot2cold" % flags)
        run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 0.01 * s, 0.002 * s, "dbg%s_4096_cold" % flags)
        run_dense(1024, capi.SWEEP_F64, 148 * 16, 8, 0.3 * 32, 0.02 * 32, "dbg%s_1024_f64" % flags)
    os.environ.pop("OSA_WS_DEBUG")

if "dense4k" in what:  # the N = 4096 fp32 one-wave probes alone (A/B runs of apply-loop variants)
    s = np.sqrt(4096)
    run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 0.3 * s, 0.02 * s, "dense4096_f32_1wave_hot2cold")
    run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 0.01 * s, 0.002 * s, "dense4096_f32_cold")
    run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 2 * s, 1 * s, "dense4096_f32_hot")

if "dens" in what:
    # apply warps alone (debug flag 2) on pseudo-random masks of density 3/16, 1/16, 1/32, 1/64 per
    # trajectory: cost of a block as a function of the rows it streams (fixed cost per block vs per row)
    import os
    s = np.sqrt(4096)
    for flags in ("2", "66", "130", "194"):
        os.environ["OSA_WS_DEBUG"] = flags
        run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 0.3 * s, 0.02 * s, "dens%s_4096" % flags)
    os.environ.pop("OSA_WS_DEBUG")

if "flowdbg" in what:
    # role experiments of the free-running kernel (probe builds): 1 = decide warps alone,
    # 2 / 66 / 130 / 194 = apply warps alone on masks of density 3/16, 1/16, 1/32, 1/64
    import os
    s = np.sqrt(4096)
    os.environ["OSA_WS_DEBUG"] = "1"
    run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 0.3 * s, 0.02 * s, "flowdbg1_hot2cold")
    run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 0.01 * s, 0.002 * s, "flowdbg1_cold")
    run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 2 * s, 1 * s, "flowdbg1_hot")
    for flags in ("2", "66", "130", "194"):
        os.environ["OSA_WS_DEBUG"] = flags
        run_dense(4096, capi.SWEEP_F32, 148 * 12, 4, 0.3 * s, 0.02 * s, "flowdbg%s" % flags)
    os.environ.pop("OSA_WS_DEBUG")

if "benchlike" in what:
    # the bench's schedule on 2 waves of trajectories
    tries = 148 * 12 * 2
    q = gen.dense_uniform_qubo(4096, seed=2029)
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as prob:
        sched = geo(32, 1.28, 19.2)
        for rep in range(2):
            res = prob.anneal(sched, 32, tries, mode=capi.MODE_SEQUENTIAL_SWEEP)
        st = res.stats
    sec = st["ms_sweep"] * 1e-3
    gbs = (st["row_fetches"] + st["init_row_fetches"]) * 16384 / sec / 1e9
    print(json.dumps({"probe": "benchlike_4096_f32", "tries": tries, "ms_sweep": round(st["ms_sweep"], 3),
                      "attempts_per_s": st["attempts"] / sec, "accept_frac": st["accepts"] / st["attempts"],
                      "row_gbs": gbs, "frac_of_18223": gbs / 18223.7,
                      "kcyc_per_cta": {k: round(st[k] / max(1, st["grid"]) / 1e3, 1)
                                       for k in ("cyc_init", "cyc_stage", "cyc_decide", "cyc_apply")}}))

if "cfg" in what:
    import os
    s = np.sqrt(4096)
    for cfg in os.environ.get("OSA_PROBE_CFGS", "821,822,824,731,631,632,541,441,451,452,411").split(","):
        os.environ["OSA_DS_CFG"] = cfg
        r = 8 if cfg[0] in "85" else int(cfg[:2])
        run_dense(4096, capi.SWEEP_F32, 148 * r, 4, 0.3 * s, 0.02 * s, "cfg%s_hot2cold" % cfg)
    os.environ.pop("OSA_DS_CFG")

if "wsr" in what:
    import os
    s = np.sqrt(4096)
    for r, key in ((8, "8"), (10, "10"), (12, "12")):
        os.environ["OSA_WS_R"] = key
        run_dense(4096, capi.SWEEP_F32, 148 * r, 4, 0.3 * s, 0.02 * s, "ws_R%s_hot2cold" % key)
    os.environ.pop("OSA_WS_R")

if "sparse" in what:
    n = 5627
    rowptr, col, val, diag = gen.sparse_random_graph(n, 15, seed=2028)
    for prec, name in ((capi.SWEEP_F32, "f32"), (capi.SWEEP_F64, "f64")):
        with Problem.csr(rowptr, col, val, diag, sweep_precision=prec) as prob:
            sched = np.linspace(0.05, 2.0, 10)
            for tries in (65536,):
                res = prob.anneal(sched, 10, tries, mode=capi.MODE_SEQUENTIAL_SWEEP)
                st = res.stats
                sec = st["ms_sweep"] * 1e-3
                print(json.dumps({"probe": "sparse5627_" + name, "tries": tries, "sweeps": 10,
                                  "grid": st["grid"], "ms_sweep": round(st["ms_sweep"], 3),
                                  "ms_energy": round(st["ms_energy"], 3),
                                  "attempts_per_s": st["attempts"] / sec,
                                  "accept_frac": st["accepts"] / st["attempts"],
                                  "energy": res.energy}))
