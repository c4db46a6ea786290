#!/usr/bin/env python
"""bench.py -- spin-flip attempts/s of the annealing hot path on B200 (BASELINE.json metric).

Workload (N=1 and per GPU at N>1, weak scaling): BASELINE config 5's per-GPU share --
synthetic dense random QUBO, N=4096, U(-1,1) coefficients, 131072 trajectories per GPU
(1M tries / 8 GPUs), sequential-sweep mode, the reference's acceptance rule
(annealing.hpp:106-108) on a geometric schedule built like one-solver-anneal.cpp:31-39.
One "step" = one full osa_anneal call on that batch: initial fields + all sweeps + exact
fp64 energies + argmin (+ the NCCL best-energy gather at N>1).

  python bench.py --gpus N --steps K --warmup W            engine arm (this repo's CUDA path)
  python bench.py --impl reference ...                     the reference's own algorithm
                                                           (oracle restatement) on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "spin-flip attempts/s, dense N=4096"
UNIT = "attempts/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    # config5 (default) is the headline workload of BASELINE.json's metric; config3 / config4 run
    # the other GPU configs of BASELINE.json through the same timing harness (not bench lines of
    # the round: they exist so that those shapes can be measured at 1/2/4/8 GPUs as well)
    ap.add_argument("--workload", default="config5",
                    choices=["config5", "config3", "config4", "random_site", "config5_f64", "config5_r8"])
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the config3 / config4 / random-site / fp64 sub-records of the headline line")
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--tries-per-gpu", type=int, default=131072)
    ap.add_argument("--sweeps", type=int, default=32)
    ap.add_argument("--beta-min", type=float, default=1.28)   # 0.02 * sqrt(N)
    ap.add_argument("--beta-max", type=float, default=19.2)   # 0.30 * sqrt(N)
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------- helpers
def sm_count(device):
    """Number of SMs of the device, asked from the CUDA runtime the engine library already loaded."""
    import ctypes
    try:
        from onesolver_b200 import capi
        capi.load()
        rt = ctypes.CDLL("libcudart.so.12")
        value = ctypes.c_int(0)
        if rt.cudaDeviceGetAttribute(ctypes.byref(value), 16, int(device)) == 0 and value.value > 0:
            return value.value  # 16 = cudaDevAttrMultiProcessorCount
    except OSError:
        pass
    return 148  # B200


def load_measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, devices):
        self.devices = list(devices) if isinstance(devices, (list, tuple)) else [devices]
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--id=" + ",".join(str(d) for d in self.devices), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_instance(n):
    from onesolver_b200 import problems as gen
    return gen.dense_uniform_qubo(n, seed=2024 + 5)  # instance seed 2024 + config number


def make_schedule(args):
    from onesolver_b200 import construct_geometric_beta_schedule
    if args.sweeps == 1:
        return np.array([args.beta_min])
    return construct_geometric_beta_schedule(args.beta_min, args.beta_max, args.sweeps)


def config_dict(args, world):
    return {"workload": f"BASELINE config 5 per-GPU share (1M tries / 8 GPUs): dense N={args.n} U(-1,1) QUBO, "
                        f"{args.tries_per_gpu} tries/GPU, {args.sweeps} sequential sweeps, "
                        f"reference accept rule exp(-dE/beta)>u, geometric beta "
                        f"{args.beta_min}->{args.beta_max}",
            "n": args.n, "tries_per_gpu": args.tries_per_gpu, "sweeps": args.sweeps,
            "mode": "sequential_sweep", "accept_rule": "reference", "schedule": "geometric",
            "beta_min": args.beta_min, "beta_max": args.beta_max, "sweep_precision": args.precision,
            "seed": 1234, "parallelism": f"trajectory shards x{world}, Q replicated",
            "l2_hygiene": "Q copy streamed by the sweep is L2-resident by design (64 MiB fp32); "
                          "each step also streams the 128 MiB fp64 copy (energy kernel) and writes "
                          "64 MiB of states, which evicts it between steps"}


# --------------------------------------------------------------------------- reference arm
def cpu_reference_sample(args, q, seconds_target=12.0):
    """Reference algorithm (full O(N^2) energy per attempt, annealing.hpp:85-126) restated in
    oracle/osa_oracle.c, all host threads, bounded sample of the same instance."""
    from oracle import binding as ob
    # all host cores, regardless of OMP_NUM_THREADS (torchrun sets it to 1)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    n = q.shape[0]
    flat = np.ascontiguousarray(q)
    tries = cores * 2
    # ~ n^2/2 MACs per attempt at ~1 GMAC/s/core
    est_attempt_s = max(1e-7, n * n / 2 / 0.8e9)
    iters = int(max(4, min(2000, seconds_target * cores / (tries * est_attempt_s))))
    sched = np.geomspace(args.beta_min, args.beta_max, iters)
    t0 = time.perf_counter()
    ob.ref_anneal(flat, n, sched, iters, tries, num_threads=cores)
    dt = time.perf_counter() - t0
    attempts = tries * iters
    return {"value": attempts / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{tries} tries x {iters} single-flip attempts (random-site, full energy "
                      f"recompute) on the same N={n} instance, {dt:.2f} s wall; includes "
                      f"{tries} initial energy evaluations"}, attempts, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    q = make_instance(args.n)
    vals, times = [], []
    base = None
    for i in range(args.warmup + args.steps):
        base, attempts, dt = cpu_reference_sample(args, q, seconds_target=4.0)
        if i >= args.warmup:
            vals.append(attempts)
            times.append(dt)
    value = sum(vals) / sum(times)
    base["value"] = value
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * sum(times) / max(1, args.steps), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(args, args.gpus), "cpu_baseline": base,
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------- engine arm
# All GPUs of the box are driven by ONE process through the product's multi-device entry
# (osa_multi_anneal: one host thread and stream per GPU, Q replicated, trajectories sharded by global
# id, one NCCL all-gather of the best records).  Under torchrun (the driver's launch for N > 1) rank 0
# makes that call over devices 0..N-1; the other ranks take part in the barriers that bracket the
# timed regions and in the max-over-ranks of the clocks, nothing else.
class Ranks:
    """torch.distributed plumbing of the bench contract (barrier + synchronize, max over ranks).

    The ranks meet on a gloo (CPU) group: an NCCL barrier posted early by an idle rank is a kernel
    that spins on its GPU until the last rank arrives, and the GPU then time-slices between that
    process and rank 0's annealing kernels on the same device (measured: 2.25x slower).  The NCCL
    process group is still created and exercised once before the timed regions."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = self.torch = self.cpu_group = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            t = torch.ones(1, device="cuda")
            dist.all_reduce(t)  # one collective over NVLink: every rank is up and sees its GPU
            torch.cuda.synchronize()
            assert int(t.item()) == self.world
            self.cpu_group = dist.new_group(backend="gloo")
            self.dist, self.torch = dist, torch
        self.active = self.rank == 0

    def barrier(self):
        if self.dist is not None:
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.cpu_group)
            self.torch.cuda.synchronize()

    def max(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.cpu_group)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.barrier(group=self.cpu_group)
            self.dist.destroy_process_group()


def nccl_log_setup():
    """NCCL's own log of the library's communicator (ncclCommInitAll over the N devices): kept on,
    in a file when the caller did not ask for it on the console."""
    if "NCCL_DEBUG" not in os.environ:
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/osa_bench_nccl_%h_%p.log")
    return os.environ.get("NCCL_DEBUG_FILE")


def nccl_log_summary(path_pattern):
    """nranks of the communicators this process created, from NCCL's log file (None: console)."""
    if not path_pattern:
        return None
    import glob
    import re
    import socket
    path = path_pattern.replace("%h", socket.gethostname()).replace("%p", str(os.getpid()))
    ranks = set()
    for f in glob.glob(path):
        try:
            for line in open(f, errors="replace"):
                m = re.search(r"nranks (\d+)", line)
                if m:
                    ranks.add(int(m.group(1)))
        except OSError:
            pass
    return sorted(ranks)


def measure(ranks, make_problem, sched, sweeps, total_tries, steps, warmup, mode, e2e_make=None):
    """W warm-up calls, then K timed calls of the multi-device anneal on the resident problem,
    bracketed by barrier + synchronize; optionally the same through host buffers (create + anneal
    + destroy per step).  Returns a dict on every rank (timings are max over ranks)."""
    out = {}
    prob = make_problem() if ranks.active else None
    agg = {"attempts": 0, "accepts": 0, "row_fetches": 0, "init_row_fetches": 0, "launches": 0,
           "ms_total": 0.0, "ms_sweep": 0.0, "ms_energy": 0.0}
    last = None
    if ranks.active:
        for _ in range(warmup):
            prob.anneal(sched, sweeps, total_tries, mode=mode)
    ranks.barrier()
    t0 = time.perf_counter()
    if ranks.active:
        for _ in range(steps):
            last = prob.anneal(sched, sweeps, total_tries, mode=mode)
            for k in agg:
                agg[k] += last.stats[k]
    ranks.barrier()
    out["wall_ms"] = ranks.max((time.perf_counter() - t0) * 1e3)
    out["agg"], out["last"] = agg, last
    if ranks.active:
        prob.close()
    if e2e_make is not None:
        def e2e_step():
            with e2e_make() as p2:  # host arrays -> device layouts on every GPU, every step
                p2.anneal(sched, sweeps, total_tries, mode=mode, want_energies=True)
        if ranks.active:
            # warm-up cycles (communicator, memory pool: the first create allocates device memory,
            # later ones reuse it -- osa_api.cu allocates the upload buffer after the resident
            # arrays so that the pool does not re-map from cycle to cycle, tools/pool_probe.cu)
            for _ in range(max(1, min(3, warmup))):
                e2e_step()
        ranks.barrier()
        t1 = time.perf_counter()
        if ranks.active:
            for _ in range(steps):
                e2e_step()
        ranks.barrier()
        out["e2e_ms"] = ranks.max((time.perf_counter() - t1) * 1e3)
    return out


def run_engine(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
               "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    nccl_log = nccl_log_setup()
    ranks = Ranks()  # (imports torch first when N > 1: one libnccl per process, see osa_multi.cu)

    from onesolver_b200 import MultiProblem, capi, measure_read_bandwidth, device_name

    devices = list(range(world))
    q = make_instance(args.n)
    sched = make_schedule(args)
    prec = capi.SWEEP_F32 if args.precision == "f32" else capi.SWEEP_F64
    esz = 4 if args.precision == "f32" else 8
    tries = args.tries_per_gpu * world  # weak scaling: every GPU anneals tries_per_gpu trajectories
    mode = capi.MODE_SEQUENTIAL_SWEEP

    sampler = ClockSampler(devices)
    if ranks.active:
        sampler.start()
    e2e_make = None
    free_pinned = None
    if not args.no_e2e and ranks.active:
        from onesolver_b200 import pinned_copy
        q_pinned, free_pinned = pinned_copy(q)  # the step's input lives in pinned host memory
        e2e_make = lambda: MultiProblem.dense(q_pinned, devices=devices, sweep_precision=prec)  # noqa: E731
    elif not args.no_e2e:
        e2e_make = lambda: None  # noqa: E731  (inactive ranks only keep the barriers in step)
    m = measure(ranks, lambda: MultiProblem.dense(q, devices=devices, sweep_precision=prec), sched,
                args.sweeps, tries, args.steps, args.warmup, mode, e2e_make)
    clocks = sampler.stop() if ranks.active else None
    if free_pinned:
        free_pinned()
    # the same through a pageable std::vector-like host buffer (what a caller of sa::anneal has);
    # a few steps are enough for this side figure
    pageable = None
    pageable_steps = min(args.steps, 3)
    if not args.no_e2e and world == 1:
        t1 = time.perf_counter()
        for _ in range(pageable_steps):
            with MultiProblem.dense(q, devices=devices, sweep_precision=prec) as p2:
                p2.anneal(sched, args.sweeps, tries, mode=mode, want_energies=True)
        pageable = (time.perf_counter() - t1) * 1e3

    others = []
    if not args.no_other_configs:
        for name in ("config3", "config4", "random_site", "config5_f64", "config5_r8"):
            others.append(run_sub_record(ranks, args, name, devices))

    if ranks.active:
        agg, last = m["agg"], m["last"]
        st = last.stats
        attempts_per_step = st["attempts"]
        # N = 1: device time of the call (CUDA events on the library's stream); N > 1: wall clock of
        # the barrier-bracketed region (the devices run concurrently on their own streams)
        step_ms = (m["wall_ms"] if world > 1 else agg["ms_total"]) / args.steps
        value = attempts_per_step / (step_ms * 1e-3)
        ld = -(-args.n // (1024 if esz == 4 else 512)) * (1024 if esz == 4 else 512)
        row_bytes = ld * esz
        # dominant kernel: the sweep kernel (init fields + sweeps), one launch per step and device;
        # bytes and time per device (the devices run the same kernel side by side)
        alg_bytes_per_launch = (agg["row_fetches"] + agg["init_row_fetches"]) * row_bytes / args.steps / world
        sweep_s_per_launch = agg["ms_sweep"] / args.steps * 1e-3   # slowest device
        achieved = alg_bytes_per_launch / sweep_s_per_launch / 1e9
        # best of five: the probe's result moves by +-8% from call to call, the peak is its maximum
        l2_peak = max(measure_read_bandwidth(64 << 20, 64, device=0) for _ in range(5))
        peaks, peaks_src = load_measured_peaks()
        q_bytes = args.n * ld * esz
        bound = "l2" if q_bytes <= 100 * (1 << 20) else "hbm"
        peak = l2_peak if bound == "l2" else peaks["hbm_gbs"]
        default_workload = (args.n == 4096 and args.tries_per_gpu == 131072 and args.sweeps == 32
                            and args.precision == "f32" and args.beta_min == 1.28
                            and args.beta_max == 19.2)
        traffic, traffic_note = load_traffic(default_workload)
        # exact-energy kernel: 16 DMMA (m8n8k4 = 512 FLOP) per k-step, k up to the diagonal block
        nblk = (args.n + 31) // 32
        dmma_per_tile = 16 * sum((32 * b + 32) // 4 for b in range(nblk))
        energy_flops = ((args.tries_per_gpu + 31) // 32) * dmma_per_tile * 512.0
        energy_tflops = energy_flops / (agg["ms_energy"] / args.steps * 1e-3) / 1e12
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        dmma_peak = 128.0 * sm_count(0) * sm_mhz * 1e6 / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": config_dict(args, world),
            "device": device_name(0),
            "entry": "osa_multi_anneal (include/onesolver_b200.h): one process, one host thread and "
                     "stream per GPU, one ncclAllGather of the best records",
            # comm_devices: size of the communicator the library built (ncclCommInitAll over the
            # devices of the osa_multi handle; 1 device = no communicator); comm_nranks_seen: the
            # same number as NCCL's own INFO log states it, when that log went to a file
            "nccl": {"comm_devices": int(st.get("reserved", 0)) or 1,
                     "comm_nranks_seen": nccl_log_summary(nccl_log),
                     "log": nccl_log or ("console (NCCL_DEBUG=%s set by the caller)"
                                         % os.environ.get("NCCL_DEBUG", "?"))},
            "breakdown_ms_per_step": {"sweep_kernel": agg["ms_sweep"] / args.steps,
                                      "exact_energy_kernel": agg["ms_energy"] / args.steps,
                                      "device_total": agg["ms_total"] / args.steps,
                                      "wall": m["wall_ms"] / args.steps},
            "accept_frac": agg["accepts"] / max(1, agg["attempts"]),
            "traj_per_row_fetch": st["traj_per_batch"],
            "sweep_only_attempts_per_s": attempts_per_step / (agg["ms_sweep"] / args.steps * 1e-3),
            "roofline": {
                "bound": bound, "kernel": "k_dense_seq (init fields + sweeps), per device",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": ("live osa_measure_read_bandwidth over a 64 MiB L2-resident buffer, best of 5"
                                if bound == "l2" else f"MEASURED_PEAKS.json hbm_gbs ({peaks_src})"),
                "hbm_peak": peaks["hbm_gbs"], "hbm_peak_source": peaks_src,
                "frac_of_hbm_peak": achieved / peaks["hbm_gbs"],
                "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                "bytes_unshared_per_launch": agg["accepts"] * row_bytes / args.steps / world,
                "traffic": traffic, "traffic_unit": "bytes per launch",
                "traffic_note": traffic_note,
            },
            "energy_kernel": {
                "kernel": "k_energy_dense_mma (FP64 tensor cores, DMMA m8n8k4)",
                "bound": "tensor", "achieved": energy_tflops, "peak": dmma_peak, "unit": "TFLOP/s",
                "frac": energy_tflops / dmma_peak,
                "flops_per_launch": energy_flops,
                "peak_source": "128 fp64 tensor FLOP/clk/SM (ncu sm__ops_path_tensor_src_fp64 "
                               "peak_sustained) x SMs x the SM clock sampled during the run",
            },
            "clocks": clocks,
            "gpu_launches": agg["launches"],
            "best_energy": last.energy, "best_index": last.index,
        }
        if "e2e_ms" in m:
            h2d = (args.n * args.n * 8 + args.sweeps * 8) * world
            d2h = tries * 8 + (((args.n + 31) // 32) * 4 + 16) * world * world + 64
            out["e2e"] = {"value": attempts_per_step / (m["e2e_ms"] / args.steps * 1e-3), "unit": UNIT,
                          "ms_per_step": m["e2e_ms"] / args.steps,
                          "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                          "what": "osa_multi_create_dense_f64(pinned host Q) + osa_multi_anneal(host "
                                  "outputs) + osa_multi_destroy per step, wall clock"}
            if pageable is not None:
                out["e2e"]["pageable_input"] = {
                    "value": attempts_per_step / (pageable / pageable_steps * 1e-3),
                    "ms_per_step": pageable / pageable_steps, "steps": pageable_steps,
                    "what": "the same with Q in ordinary (pageable) host memory, as a caller of "
                            "sa::anneal has it"}
        if others:
            out["other_configs"] = [o for o in others if o]
        if not args.no_cpu_baseline and world == 1:
            base, _, _ = cpu_reference_sample(args, q)
            out["cpu_baseline"] = base
        print(json.dumps(out), flush=True)
    ranks.close()


def load_traffic(default_workload):
    """DRAM traffic of one sweep launch at the default workload, from the committed ncu capture."""
    if not default_workload:
        return None, "not captured for this workload"
    path = os.path.join(ROOT, "profiles", "r02", "ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        return t["dram_bytes"], t["note"]
    except Exception:
        return 136361905408 + 1937762304, (
            "dram__bytes_read.sum + dram__bytes_write.sum of one launch at this workload (ncu, "
            "profiles/r01/ncu_traffic_v54.csv): 0.14 TB of DRAM traffic against 14.08 TB of "
            "algorithmic row bytes, which are served by the L2")


# --------------------------------------------------------------------------- configs 3 and 4
def other_config_spec(args):
    """BASELINE.json configs 3 (dense fp64 N=1024, 16384 tries, 1000 sweeps) and 4 (sparse
    Pegasus-like N=5627 CSR, 65536 tries, linear schedule), and the reference's own random-site loop
    on the headline instance: instance, schedule, problem factory.
    Weak scaling: every GPU anneals the config's full number of tries."""
    from onesolver_b200 import (MultiProblem, capi, construct_geometric_beta_schedule,
                                construct_linear_beta_schedule)
    from onesolver_b200 import problems as gen
    if args.workload == "config3":
        n, tries, sweeps = 1024, 16384, 1000
        q = gen.dense_uniform_qubo(n, seed=2024 + 3)
        sched = construct_geometric_beta_schedule(0.02 * 32, 0.30 * 32, sweeps)
        return {"metric": "spin-flip attempts/s, dense fp64 N=1024", "n": n, "tries": tries,
                "sweeps": sweeps, "sched": sched, "dtype": "f64", "esz": 8,
                "mode": capi.MODE_SEQUENTIAL_SWEEP,
                "make": lambda devs, src=q: MultiProblem.dense(src, devices=devs, sweep_precision=capi.SWEEP_F64),
                "host_input": q, "h2d": q.nbytes + sched.nbytes,
                "workload": f"BASELINE config 3: dense fp64 N={n} U(-1,1) QUBO, {tries} tries/GPU, "
                            f"{sweeps} sequential sweeps, reference accept rule, geometric beta 0.64->9.6"}
    if args.workload == "config5_f64":
        # SURVEY 8(d), C5 secondary: the headline instance and schedule with fp64 fields.  The fp64
        # copy of Q is 128 MiB -- just above the 126 MB L2 -- and a thread can hold the fields of
        # 4 trajectories only (R = 4), so a row fetch is shared by 4 instead of 12.
        n, tries, sweeps = 4096, 16384, 32
        q = gen.dense_uniform_qubo(n, seed=2024 + 5)
        sched = construct_geometric_beta_schedule(1.28, 19.2, sweeps)
        return {"metric": "spin-flip attempts/s, dense N=4096, fp64 fields", "n": n, "tries": tries,
                "sweeps": sweeps, "sched": sched, "dtype": "f64", "esz": 8,
                "mode": capi.MODE_SEQUENTIAL_SWEEP,
                "make": lambda devs, src=q: MultiProblem.dense(src, devices=devs, sweep_precision=capi.SWEEP_F64),
                "host_input": q, "h2d": q.nbytes + sched.nbytes,
                "note": "the streamed copy of Q (128 MiB) does not fit the L2 entirely: compare "
                        "roofline.achieved with MEASURED_PEAKS.json hbm_gbs as well",
                "workload": f"BASELINE config 5 instance with fp64 fields: dense N={n}, {tries} tries/GPU, "
                            f"{sweeps} sequential sweeps, reference accept rule, geometric beta 1.28->19.2"}
    if args.workload == "config5_r8":
        # the headline workload on fewer trajectories, with R = 8 instead of 12 trajectories sharing a
        # row fetch (OSA_FLOW_R, a result-preserving knob of the free-running kernel): the rows go by
        # faster (higher roofline fraction) and each serves fewer attempts (lower throughput) --
        # the trade-off behind the headline line's roofline.frac, measured by the driver as well
        n, tries, sweeps = 4096, 32768, 32
        q = gen.dense_uniform_qubo(n, seed=2024 + 5)
        sched = construct_geometric_beta_schedule(1.28, 19.2, sweeps)
        return {"metric": "spin-flip attempts/s, dense N=4096, 8 trajectories per row fetch", "n": n,
                "tries": tries, "sweeps": sweeps, "sched": sched, "dtype": "f32", "esz": 4,
                "mode": capi.MODE_SEQUENTIAL_SWEEP, "env": {"OSA_FLOW_R": "8"},
                "make": lambda devs, src=q: MultiProblem.dense(src, devices=devs, sweep_precision=capi.SWEEP_F32),
                "host_input": q, "h2d": q.nbytes + sched.nbytes,
                "workload": f"headline instance and schedule, {tries} tries/GPU, R = 8 trajectories per "
                            f"CTA instead of 12 (OSA_FLOW_R=8): roofline fraction against throughput"}
    if args.workload == "random_site":
        # the reference's loop (annealing.hpp:97-101): one attempt per iteration at a random site
        n, tries, iters = 4096, 16384, 4096
        q = gen.dense_uniform_qubo(n, seed=2024 + 5)
        sched = construct_geometric_beta_schedule(1.28, 19.2, iters)
        return {"metric": "spin-flip attempts/s, dense N=4096, random-site (reference loop)", "n": n,
                "tries": tries, "sweeps": iters, "sched": sched, "dtype": "f32", "esz": 4,
                "mode": capi.MODE_RANDOM_SITE,
                "make": lambda devs, src=q: MultiProblem.dense(src, devices=devs, sweep_precision=capi.SWEEP_F32),
                "host_input": q, "h2d": q.nbytes + sched.nbytes,
                "workload": f"reference loop on the headline instance: dense N={n}, {tries} tries/GPU, "
                            f"{iters} single-flip attempts per trajectory at random sites (annealing.hpp:"
                            f"97-101), reference accept rule, geometric beta 1.28->19.2, fp32 fields"}
    n, tries, sweeps = 5627, 65536, 100
    rowptr, col, val, diag = gen.sparse_random_graph(n, 15, seed=2024 + 4)
    sched = construct_linear_beta_schedule(0.5, 5.0, sweeps)
    prec = capi.SWEEP_F32 if args.precision == "f32" else capi.SWEEP_F64
    return {"metric": "spin-flip attempts/s, sparse N=5627 (CSR)", "n": n, "tries": tries,
            "sweeps": sweeps, "sched": sched, "dtype": args.precision,
            "esz": 4 if args.precision == "f32" else 8, "nnz": int(len(col)),
            "mode": capi.MODE_SEQUENTIAL_SWEEP,
            "make": lambda devs: MultiProblem.csr(rowptr, col, val, diag, devices=devs, sweep_precision=prec),
            "host_input": None,
            "h2d": rowptr.nbytes + col.nbytes + val.nbytes + diag.nbytes + sched.nbytes,
            "workload": f"BASELINE config 4: sparse Pegasus-like N={n} CSR ({len(col) // 2} couplers, "
                        f"degree <= 15), {tries} tries/GPU, {sweeps} sequential sweeps, reference "
                        f"accept rule, linear beta 0.5->5.0, fields in {args.precision}"}


def sub_record(spec, m, world, steps, warmup, l2_peak, clocks):
    """One record of the bench line for a workload other than the headline one."""
    agg, st = m["agg"], m["last"].stats
    attempts_per_step = st["attempts"]
    step_ms = (m["wall_ms"] if world > 1 else agg["ms_total"]) / steps
    sweep_s = agg["ms_sweep"] / steps * 1e-3
    if "nnz" in spec:
        # every warp (32 trajectories) streams the CSR once per sweep: nnz x (index + value)
        alg = -(-spec["tries"] // 32) * spec["sweeps"] * spec["nnz"] * (4 + spec["esz"])
        what = "CSR blocks streamed per device: warps x sweeps x nnz x (4 B index + value)"
    else:
        unit = 1024 if spec["esz"] == 4 else 512
        ld = -(-spec["n"] // unit) * unit
        alg = (agg["row_fetches"] + agg["init_row_fetches"]) * ld * spec["esz"] / steps / world
        what = "Q rows streamed per device: (row_fetches + init_row_fetches) x ld x element size"
    rec = {"metric": spec["metric"], "value": attempts_per_step / (step_ms * 1e-3), "unit": UNIT,
           "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": step_ms,
           "scaling": "weak", "dtype": spec["dtype"],
           "config": {"workload": spec["workload"], "n": spec["n"], "tries_per_gpu": spec["tries"],
                      "sweeps": spec["sweeps"], "accept_rule": "reference", "seed": 1234},
           "breakdown_ms_per_step": {"sweep_kernel": agg["ms_sweep"] / steps,
                                     "exact_energy_kernel": agg["ms_energy"] / steps,
                                     "device_total": agg["ms_total"] / steps, "wall": m["wall_ms"] / steps},
           "accept_frac": agg["accepts"] / max(1, agg["attempts"]),
           "kernel": st["kernel_id"], "traj_per_row_fetch": st["traj_per_batch"],
           "roofline": {"bound": "l2", "kernel": "sweep kernel, per device",
                        "achieved": alg / sweep_s / 1e9, "peak": l2_peak, "unit": "GB/s",
                        "frac": alg / sweep_s / 1e9 / l2_peak,
                        "algorithmic_bytes_per_launch": alg, "algorithmic_bytes": what, "traffic": None},
           "clocks": clocks, "gpu_launches": agg["launches"],
           "best_energy": m["last"].energy, "best_index": m["last"].index}
    if "note" in spec:
        peaks, _ = load_measured_peaks()
        rec["roofline"]["hbm_peak"] = peaks.get("hbm_gbs")
        rec["roofline"]["frac_of_hbm_peak"] = alg / sweep_s / 1e9 / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None
        rec["roofline"]["note"] = spec["note"]
    if "e2e_ms" in m:
        rec["e2e"] = {"value": attempts_per_step / (m["e2e_ms"] / steps * 1e-3), "unit": UNIT,
                      "ms_per_step": m["e2e_ms"] / steps, "h2d_bytes_per_step": int(spec["h2d"]) * world,
                      "d2h_bytes_per_step": spec["tries"] * world * 8 + ((spec["n"] + 31) // 32) * 4 + 16 + 64,
                      "what": "osa_multi_create_*(host arrays, the dense matrix in pinned memory) + "
                              "osa_multi_anneal(host outputs) + osa_multi_destroy per step, wall clock"}
    return rec


SUB_RECORD_STEPS = {"config3": 1, "config4": 2, "random_site": 5, "config5_f64": 2, "config5_r8": 2}


def run_sub_record(ranks, args, name, devices):
    """A reduced-step measurement of another workload inside the headline run: W = 3 warm-up
    steps; K = 1 for the 1.5 s step of config 3, 2 for the 0.4-0.6 s steps, 5 for the 55 ms step of
    the random-site loop (its host-side share of the end-to-end step is not a single sample then)."""
    import copy
    a = copy.copy(args)
    a.workload = name
    world = len(devices)
    nsteps, nwarm = SUB_RECORD_STEPS[name], 3
    spec = other_config_spec(a) if ranks.active else None
    sampler = ClockSampler(devices)
    if ranks.active:
        sampler.start()
        make = lambda: spec["make"](devices)  # noqa: E731
        e2e_make, free_pinned = None, None
        if not args.no_e2e:
            e2e_make = make
            if spec.get("host_input") is not None:
                # like the headline: the step's input lives in pinned host memory
                from onesolver_b200 import pinned_copy
                q_pinned, free_pinned = pinned_copy(spec["host_input"])
                e2e_make = lambda: spec["make"](devices, q_pinned)  # noqa: E731
        os.environ.update(spec.get("env", {}))
        try:
            m = measure(ranks, make, spec["sched"], spec["sweeps"], spec["tries"] * world, nsteps, nwarm,
                        spec["mode"], e2e_make)
        finally:
            for k in spec.get("env", {}):
                os.environ.pop(k, None)
            if free_pinned:
                free_pinned()
        clocks = sampler.stop()
        from onesolver_b200 import measure_read_bandwidth
        l2_peak = max(measure_read_bandwidth(64 << 20, 64, device=0) for _ in range(3))
        return sub_record(spec, m, world, nsteps, nwarm, l2_peak, clocks)
    measure(ranks, None, None, 0, 0, nsteps, nwarm, 0, None if args.no_e2e else (lambda: None))
    return None


def run_other_config(args):
    """`--workload config3|config4|random_site`: that workload alone, with the full K / W."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable":
                              "the reference arm is defined for the headline workload (config5) only"}))
        return
    if args.gpus > 1 and world == 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
               "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    nccl_log_setup()
    ranks = Ranks()
    from onesolver_b200 import measure_read_bandwidth
    devices = list(range(world))
    spec = other_config_spec(args) if ranks.active else None
    sampler = ClockSampler(devices)
    if ranks.active:
        sampler.start()
        make = lambda: spec["make"](devices)  # noqa: E731
        e2e_make, free_pinned = None, None
        if not args.no_e2e:
            e2e_make = make
            if spec.get("host_input") is not None:  # pinned host input, as in run_sub_record
                from onesolver_b200 import pinned_copy
                q_pinned, free_pinned = pinned_copy(spec["host_input"])
                e2e_make = lambda: spec["make"](devices, q_pinned)  # noqa: E731
        os.environ.update(spec.get("env", {}))
        m = measure(ranks, make, spec["sched"], spec["sweeps"], spec["tries"] * world, args.steps,
                    args.warmup, spec["mode"], e2e_make)
        if free_pinned:
            free_pinned()
        clocks = sampler.stop()
        l2_peak = max(measure_read_bandwidth(64 << 20, 64, device=0) for _ in range(5))
        rec = sub_record(spec, m, world, args.steps, args.warmup, l2_peak, clocks)
        rec.update({"higher_is_better": True, "vs_baseline": None, "data": "synthetic"})
        print(json.dumps(rec), flush=True)
    else:
        measure(ranks, None, None, 0, 0, args.steps, args.warmup, 0,
                None if args.no_e2e else (lambda: None))
    ranks.close()


def main():
    args = parse_args()
    if args.workload != "config5":
        run_other_config(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
