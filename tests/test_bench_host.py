"""CPU tests of bench.py's host logic: argument contract and the workload specifications of the
three GPU configurations of BASELINE.json (no compute calls)."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _args(**kw):
    base = dict(gpus=1, steps=1, warmup=0, impl="engine", workload="config5", n=4096,
                tries_per_gpu=131072, sweeps=32, beta_min=1.28, beta_max=19.2, precision="f32",
                no_cpu_baseline=True, no_e2e=True)
    base.update(kw)
    return argparse.Namespace(**base)


def test_default_workload_is_baseline_config5_share():
    cfg = bench.config_dict(_args(), 8)
    assert cfg["n"] == 4096 and cfg["tries_per_gpu"] == 131072 and 8 * cfg["tries_per_gpu"] == 1 << 20
    assert "config 5" in cfg["workload"] and cfg["parallelism"].startswith("trajectory shards x8")
    sched = bench.make_schedule(_args())
    assert len(sched) == 32 and np.isclose(sched[0], 1.28)


def test_headline_workload_constants_are_frozen(monkeypatch):
    """attempts/s depends on the schedule (accept fraction), and BASELINE.json fixes
    neither sweeps nor beta -- the round-1 choice is the contract from here on.  Any change of these
    defaults voids the round-over-round comparison, so it has to fail a test first."""
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse_args()
    assert (a.n, a.tries_per_gpu, a.sweeps, a.precision) == (4096, 131072, 32, "f32")
    assert (a.beta_min, a.beta_max, a.workload, a.impl) == (1.28, 19.2, "config5", "engine")
    assert a.warmup >= 3 and a.gpus == 1
    sched = bench.make_schedule(a)
    assert len(sched) == 32 and sched[0] == 1.28 and abs(sched[-1] - 19.2) < 1e-9
    ratio = sched[1:] / sched[:-1]
    assert np.allclose(ratio, ratio[0])  # geometric, one-solver-anneal.cpp:31-39
    cfg = bench.config_dict(a, 1)
    assert cfg["mode"] == "sequential_sweep" and cfg["accept_rule"] == "reference"
    assert cfg["seed"] == 1234 and cfg["sweep_precision"] == "f32"


def test_config3_and_config4_specs_match_baseline_shapes():
    s3 = bench.other_config_spec(_args(workload="config3"))
    assert (s3["n"], s3["tries"], s3["sweeps"], s3["dtype"]) == (1024, 16384, 1000, "f64")
    assert s3["host_input"].shape == (1024, 1024) and len(s3["sched"]) == 1000
    assert np.allclose(s3["host_input"], s3["host_input"].T)
    s4 = bench.other_config_spec(_args(workload="config4"))
    assert (s4["n"], s4["tries"], s4["sweeps"]) == (5627, 65536, 100)
    assert s4["nnz"] % 2 == 0 and 35000 <= s4["nnz"] // 2 <= 45000  # ~40k couplers, SURVEY 8(d)
    d = np.diff(s4["sched"])
    assert np.allclose(d, d[0])  # linear schedule
    # the headline instance with fp64 fields (SURVEY 8d, C5 secondary): same instance and schedule
    s5 = bench.other_config_spec(_args(workload="config5_f64"))
    assert (s5["n"], s5["sweeps"], s5["dtype"], s5["esz"]) == (4096, 32, "f64", 8)
    assert np.array_equal(s5["host_input"], bench.make_instance(4096))
    assert np.allclose(s5["sched"], bench.make_schedule(_args()))


def test_reference_arm_of_other_configs_reports_unavailable():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--workload", "config4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" in line


# ---------------------------------------------------------------- committed bench lines
ENGINE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
               "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "roofline",
               "clocks", "gpu_launches", "e2e"}


def _last_json_line(path):
    with open(path) as f:
        return json.loads([ln for ln in f if ln.startswith("{")][-1])


def test_committed_round_bench_lines_carry_the_contract_keys():
    """The bench lines kept under profiles/ (what DESIGN.md quotes) have every key of the bench
    contract: the headline line with roofline / e2e / cpu_baseline, the reference arm, and the
    multi-GPU lines with a whole-job value."""
    prof = os.path.join(ROOT, "profiles", "r01")
    one = _last_json_line(os.path.join(prof, "bench_v52.json"))
    assert ENGINE_KEYS <= set(one) and "cpu_baseline" in one
    assert one["n_gpus"] == 1 and one["warmup"] >= 3 and one["config"]["workload"]
    rf = one["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf)
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(one["e2e"])
    assert one["e2e"]["h2d_bytes_per_step"] >= 4096 * 4096 * 8  # the matrix is uploaded every step
    assert {"value", "unit", "cores", "kind", "sample"} <= set(one["cpu_baseline"])
    assert not set(one["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    ref = _last_json_line(os.path.join(prof, "bench_ref_v52.json"))
    assert ref["impl"] == "reference" and ref["metric"] == one["metric"] and ref["unit"] == one["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] in ("port", "reference")
    for name, n in (("bench_v55_2gpu.json", 2), ("bench_v47_4gpu.json", 4), ("bench_v58_8gpu.json", 8)):
        line = _last_json_line(os.path.join(prof, name))
        assert ENGINE_KEYS <= set(line) and line["n_gpus"] == n and line["scaling"] == "weak"
        # whole-job aggregate: close to n times the one-GPU value of the same round
        assert 0.9 * n * 8.9e9 < line["value"] < 1.1 * n * one["value"]


def test_sub_records_pass_their_step_counts(monkeypatch):
    """Every sub-record of the headline line has a step count, and run_sub_record hands K / W to the
    measurement and to the record unchanged (the env-variable loop of a spec once overwrote K)."""
    import onesolver_b200
    names = ("config3", "config4", "random_site", "config5_f64", "config5_r8")
    assert set(bench.SUB_RECORD_STEPS) == set(names)
    seen = {}

    class Ranks:
        active = True

    class Sampler:
        def __init__(self, devices):
            pass

        def start(self):
            pass

        def stop(self):
            return {"sm_mhz": 0}

    def fake_spec(a):
        return {"make": lambda devs, src=None: ("problem", src), "sched": [1.0], "sweeps": 1, "tries": 7,
                "mode": 0, "env": {"OSA_FLOW_R": "8"}, "host_input": np.zeros((2, 2))}

    def fake_measure(ranks, make, sched, sweeps, tries, steps, warmup, mode, e2e_make=None):
        seen["measure"] = (steps, warmup, tries)
        seen["env_inside"] = os.environ.get("OSA_FLOW_R")
        seen["e2e_src"] = e2e_make()[1] if e2e_make else None
        return {"m": 1}

    def fake_record(spec, m, world, steps, warmup, l2_peak, clocks):
        return {"steps": steps, "warmup": warmup}

    monkeypatch.setattr(bench, "other_config_spec", fake_spec)
    monkeypatch.setattr(bench, "measure", fake_measure)
    monkeypatch.setattr(bench, "sub_record", fake_record)
    monkeypatch.setattr(bench, "ClockSampler", Sampler)
    monkeypatch.setattr(onesolver_b200, "measure_read_bandwidth", lambda *a, **k: 1.0)
    monkeypatch.setattr(onesolver_b200, "pinned_copy", lambda q: ("pinned", lambda: seen.setdefault("freed", True)))
    monkeypatch.delenv("OSA_FLOW_R", raising=False)
    for name in names:
        rec = bench.run_sub_record(Ranks(), _args(no_e2e=False), name, [0, 1])
        k = bench.SUB_RECORD_STEPS[name]
        assert rec == {"steps": k, "warmup": 3} and seen["measure"] == (k, 3, 14)
        assert seen["env_inside"] == "8" and "OSA_FLOW_R" not in os.environ
        assert seen["e2e_src"] == "pinned" and seen.pop("freed")
