"""GPU tests of the multi-device entry points (osa_multi_*, include/onesolver_b200.h): the call
sharded over the GPUs of one box returns bit for bit what the one-device call returns
(trajectories are keyed by global ids, the only exchange is one NCCL all-gather of
the best records, ties go to the lowest global id -- std::min_element, annealing.hpp:134).
Tests that need two devices skip on a one-GPU box (run them with `gpurun --gpus 2`)."""
import os
import subprocess

import numpy as np
import pytest

from onesolver_b200 import MultiProblem, Problem, capi, device_count
from onesolver_b200 import problems as gen
from oracle import binding as ob

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def geo(num_iter, lo, hi):
    return ob.ref_schedule("geometric", lo, hi, num_iter)


def same_result(a, b):
    np.testing.assert_array_equal(a.best_energies, b.best_energies)
    np.testing.assert_array_equal(a.best_states_packed, b.best_states_packed)
    np.testing.assert_array_equal(a.state, b.state)
    assert (a.energy, a.index) == (b.energy, b.index)
    for key in ("attempts", "accepts"):
        assert a.stats[key] == b.stats[key]


def test_one_device_group_equals_the_plain_call(gpu):
    """osa_multi_anneal over a single device (no communicator) against osa_anneal."""
    n, tries = 300, 50
    q = gen.dense_uniform_qubo(n, seed=7)
    sched = geo(4, 0.3, 8.0)
    kw = dict(mode=capi.MODE_SEQUENTIAL_SWEEP, want_energies=True, want_states=True, first_try=1000)
    with Problem.dense(q, sweep_precision=capi.SWEEP_F32) as p:
        one = p.anneal(sched, 4, tries, **kw)
    with MultiProblem.dense(q, devices=[0], sweep_precision=capi.SWEEP_F32) as m:
        assert m.num_devices == 1
        grp = m.anneal(sched, 4, tries, **kw)
    same_result(one, grp)
    assert grp.stats["reserved"] == 1 and len(grp.device_stats) == 1


@pytest.mark.parametrize("kind", ["dense_f32_sweep", "dense_f64_random", "csr"])
def test_two_devices_equal_one_device_bit_for_bit(gpu, kind):
    if device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    if kind == "csr":
        n = 700
        rowptr, col, val, diag = gen.sparse_random_graph(n, 6, seed=5)
        sched = np.linspace(0.05, 2.0, 6)
        kw = dict(mode=capi.MODE_SEQUENTIAL_SWEEP, want_energies=True, want_states=True)
        tries, iters = 333, 6   # odd count: the shards differ in size
        with Problem.csr(rowptr, col, val, diag) as p:
            one = p.anneal(sched, iters, tries, **kw)
        with MultiProblem.csr(rowptr, col, val, diag, num_devices=2) as m:
            two = m.anneal(sched, iters, tries, **kw)
    else:
        n = 520
        q = gen.dense_uniform_qubo(n, seed=11)
        f32 = kind == "dense_f32_sweep"
        prec = capi.SWEEP_F32 if f32 else capi.SWEEP_F64
        mode = capi.MODE_SEQUENTIAL_SWEEP if f32 else capi.MODE_RANDOM_SITE
        iters = 5 if f32 else 400
        sched = geo(iters, 0.3, 10.0)
        tries = 1001
        kw = dict(mode=mode, want_energies=True, want_states=True, first_try=77)
        with Problem.dense(q, sweep_precision=prec) as p:
            one = p.anneal(sched, iters, tries, **kw)
        with MultiProblem.dense(q, num_devices=2, sweep_precision=prec) as m:
            assert m.num_devices == 2
            two = m.anneal(sched, iters, tries, **kw)
    same_result(one, two)
    assert two.stats["reserved"] == 2
    d0, d1 = two.device_stats
    assert d0["attempts"] + d1["attempts"] == one.stats["attempts"]
    assert d0["attempts"] >= d1["attempts"] > 0      # the remainder goes to the low device


def test_fewer_trajectories_than_devices(gpu):
    """A device without trajectories contributes an infinite energy to the gather and never wins."""
    if device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    q = gen.dense_integer_qubo(24, seed=2026)
    sched = geo(50, 0.1, 3.0)
    with Problem.dense(q) as p:
        one = p.anneal(sched, 50, 1, want_energies=True, want_states=True)
    with MultiProblem.dense(q, num_devices=2) as m:
        two = m.anneal(sched, 50, 1, want_energies=True, want_states=True)
    same_result(one, two)


def test_cli_shards_over_the_visible_gpus(gpu, tmp_path):
    """one-solver-anneal --device-type gpu uses every visible GPU; --gpu-index 0 restricts it to
    one.  Both write the same result file (sa::anneal -> osa_multi_anneal)."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "app"), "-s", "-j4"], check=True)
    exe = os.path.join(ROOT, "build/bin/one-solver-anneal")
    args = ["--input", "tests/golden/chimera512/001.qubo", "--num-iter", "30", "--num-tries", "8192",
            "--mode", "sweep", "--schedule-type", "linear", "--beta-max", "3", "--device-type", "gpu",
            "--stats"]
    outs = {}
    for label, extra in (("all", []), ("one", ["--gpu-index", "0"])):
        o = tmp_path / f"{label}.csv"
        r = subprocess.run([exe] + args + ["--output", str(o)] + extra, capture_output=True,
                           text=True, cwd=ROOT)
        assert r.returncode == 0, r.stderr
        outs[label] = (o.read_text(), r.stdout)
    assert outs["all"][0] == outs["one"][0]
    if device_count() >= 2:
        assert f"x{device_count()}" in outs["all"][1] and "Devices: " in outs["all"][1]
    r = subprocess.run([exe] + args + ["--output", str(tmp_path / "x.csv"), "--gpu-index", "0",
                                       "--num-gpus", "1"], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode != 0 and "not both" in r.stderr
