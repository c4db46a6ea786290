"""Host-side mirror of the reference's annealing interface over the C ABI.

Python is only the driver for pytest and bench.py; the reference-facing host layer
is the C++ shim in include/simulated_annealing/annealing.hpp.  Names and argument
meaning follow /root/reference/include/simulated_annealing/annealing.hpp:55-58 and
/root/reference/app/one-solver-anneal.cpp:23-39.
"""
import ctypes
import math

import numpy as np

from . import capi


def construct_linear_beta_schedule(beta_min, beta_max, num_iter):
    """one-solver-anneal.cpp:23-29 (ends at beta_min + beta_max, a reference quirk)."""
    return np.array([beta_min + beta_max * i / float(num_iter - 1) for i in range(num_iter)],
                    dtype=np.float64)


def construct_geometric_beta_schedule(beta_min, beta_max, num_iter):
    """one-solver-anneal.cpp:31-39 (iterated product, not pow(alpha, i))."""
    schedule = np.empty(num_iter, dtype=np.float64)
    schedule[0] = beta_min
    alpha = math.pow(beta_max / beta_min, 1.0 / (num_iter - 1))
    for i in range(1, num_iter):
        schedule[i] = schedule[i - 1] * alpha
    return schedule


class AnnealResult:
    def __init__(self, state, energy, index, stats, best_energies=None, best_states_packed=None,
                 trace_hash=None):
        self.state = state            # uint8[N] (qubo::Solution::state)
        self.energy = energy          # float   (qubo::Solution::energy)
        self.index = index            # global id of the winning trajectory
        self.stats = stats            # dict of osa_stats
        self.best_energies = best_energies
        self.best_states_packed = best_states_packed
        self.trace_hash = trace_hash  # uint64[num_tries]: flip trace per trajectory (want_trace)


class Problem:
    """Device-resident QUBO (dense flatten_qubo layout or CSR)."""

    def __init__(self, handle, n):
        self._h = handle
        self.n = n
        self.nw = (n + 31) // 32

    @classmethod
    def dense(cls, qsym, device=0, sweep_precision=capi.SWEEP_F64):
        lib = capi.load()
        q = np.ascontiguousarray(qsym)
        if q.ndim != 2 or q.shape[0] != q.shape[1]:
            raise ValueError("qsym must be a square matrix")
        h = ctypes.c_void_p()
        if q.dtype == np.float32:
            capi.check(lib.osa_problem_create_dense_f32(q.ctypes.data, q.shape[0], device,
                                                        ctypes.byref(h)))
        else:
            q = np.ascontiguousarray(q, dtype=np.float64)
            capi.check(lib.osa_problem_create_dense_f64(q.ctypes.data, q.shape[0], device,
                                                        sweep_precision, ctypes.byref(h)))
        return cls(h, q.shape[0])

    @classmethod
    def csr(cls, rowptr, col, val, diag, device=0, sweep_precision=capi.SWEEP_F64):
        lib = capi.load()
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        diag = np.ascontiguousarray(diag, dtype=np.float64)
        n = diag.shape[0]
        _check_csr_lengths(rowptr, col, val, n)
        h = ctypes.c_void_p()
        capi.check(lib.osa_problem_create_csr_f64(rowptr.ctypes.data, col.ctypes.data,
                                                  val.ctypes.data, diag.ctypes.data, n, device,
                                                  sweep_precision, ctypes.byref(h)))
        return cls(h, n)

    def close(self):
        if self._h:
            capi.load().osa_problem_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def anneal(self, beta_schedule, num_iter, num_tries, sweeps_per_beta=1, seed=1234,
               first_try=0, mode=capi.MODE_RANDOM_SITE, accept_rule=capi.ACCEPT_REFERENCE,
               kernel_variant=capi.KID_AUTO, want_energies=False, want_states=False,
               want_trace=False):
        """sa::anneal(instance, q, beta_schedule, num_iter, num_tries, sweeps_per_beta)."""
        lib = capi.load()
        sched = np.ascontiguousarray(beta_schedule, dtype=np.float64)
        if sched.shape[0] < num_iter:
            raise ValueError("beta_schedule shorter than num_iter")
        prm = capi.AnnealParams(seed=seed, first_try=first_try, num_tries=num_tries,
                                num_iter=num_iter, sweeps_per_beta=sweeps_per_beta, mode=mode,
                                accept_rule=accept_rule, kernel_variant=kernel_variant, flags=0)
        energies = np.empty(num_tries, dtype=np.float64) if want_energies else None
        states = np.empty((num_tries, self.nw), dtype=np.uint32) if want_states else None
        trace = np.empty(num_tries, dtype=np.uint64) if want_trace else None
        state = np.empty(self.n, dtype=np.uint8)
        e = ctypes.c_double()
        idx = ctypes.c_uint64()
        st = capi.Stats()
        capi.check(lib.osa_anneal_traced(self._h, sched.ctypes.data, ctypes.byref(prm),
                                         energies.ctypes.data if want_energies else None,
                                         states.ctypes.data if want_states else None,
                                         state.ctypes.data, ctypes.byref(e), ctypes.byref(idx),
                                         trace.ctypes.data if want_trace else None,
                                         ctypes.byref(st)))
        return AnnealResult(state, e.value, idx.value, st.as_dict(), energies, states, trace)

    def parallel_tempering(self, betas, num_groups, num_rounds, sweeps_per_round, seed=1234,
                           first_group=0, accept_rule=capi.ACCEPT_BOLTZMANN, want_energies=False,
                           want_states=False):
        """osa_pt_anneal: num_groups independent ladders of len(betas) replicas (see the header)."""
        lib = capi.load()
        ladder = np.ascontiguousarray(betas, dtype=np.float64)
        m = int(ladder.shape[0])
        tries = int(num_groups) * m
        prm = capi.PtParams(seed=seed, first_group=first_group, num_groups=num_groups,
                            num_replicas=m, num_rounds=num_rounds,
                            sweeps_per_round=sweeps_per_round, accept_rule=accept_rule, flags=0,
                            reserved=0)
        energies = np.empty(tries, dtype=np.float64) if want_energies else None
        states = np.empty((tries, self.nw), dtype=np.uint32) if want_states else None
        state = np.empty(self.n, dtype=np.uint8)
        e = ctypes.c_double()
        idx = ctypes.c_uint64()
        st = capi.Stats()
        capi.check(lib.osa_pt_anneal(self._h, ladder.ctypes.data, ctypes.byref(prm),
                                     energies.ctypes.data if want_energies else None,
                                     states.ctypes.data if want_states else None,
                                     state.ctypes.data, ctypes.byref(e), ctypes.byref(idx),
                                     ctypes.byref(st)))
        return AnnealResult(state, e.value, idx.value, st.as_dict(), energies, states)

    def population_annealing(self, betas, num_populations, population_size, sweeps_per_step,
                             seed=1234, first_population=0, accept_rule=capi.ACCEPT_BOLTZMANN,
                             want_energies=False, want_states=False):
        """osa_pa_anneal: num_populations independent populations of population_size replicas
        annealed along betas with resampling between the temperatures (see the header)."""
        lib = capi.load()
        sched = np.ascontiguousarray(betas, dtype=np.float64)
        tries = int(num_populations) * int(population_size)
        prm = capi.PaParams(seed=seed, first_population=first_population,
                            num_populations=num_populations, population_size=population_size,
                            num_steps=int(sched.shape[0]), sweeps_per_step=sweeps_per_step,
                            accept_rule=accept_rule, flags=0, reserved=0)
        energies = np.empty(tries, dtype=np.float64) if want_energies else None
        states = np.empty((tries, self.nw), dtype=np.uint32) if want_states else None
        state = np.empty(self.n, dtype=np.uint8)
        e = ctypes.c_double()
        idx = ctypes.c_uint64()
        st = capi.Stats()
        capi.check(lib.osa_pa_anneal(self._h, sched.ctypes.data, ctypes.byref(prm),
                                     energies.ctypes.data if want_energies else None,
                                     states.ctypes.data if want_states else None,
                                     state.ctypes.data, ctypes.byref(e), ctypes.byref(idx),
                                     ctypes.byref(st)))
        return AnnealResult(state, e.value, idx.value, st.as_dict(), energies, states)

    def energy_batch(self, states_packed):
        """sa::energy (annealing.hpp:31-40) of packed states, evaluated on the device."""
        lib = capi.load()
        s = np.ascontiguousarray(states_packed, dtype=np.uint32)
        if s.ndim == 1:
            s = s.reshape(1, -1)
        if s.shape[1] != self.nw:
            raise ValueError("states_packed must have ceil(N/32) words per state")
        out = np.empty(s.shape[0], dtype=np.float64)
        capi.check(lib.osa_energy_batch(self._h, s.ctypes.data, s.shape[0], out.ctypes.data))
        return out


def _check_csr_lengths(rowptr, col, val, n):
    """The C side reads rowptr[0..n] and col/val[0..rowptr[n]): check the array lengths first."""
    if rowptr.ndim != 1 or rowptr.shape[0] != n + 1:
        raise ValueError(f"rowptr must have n + 1 = {n + 1} entries (got {rowptr.shape})")
    nnz = int(rowptr[n])
    if nnz < 0 or col.shape[0] < nnz or val.shape[0] < nnz:
        raise ValueError(f"col/val shorter than rowptr[n] = {nnz}")


class MultiProblem:
    """Q replicated on several GPUs of one box (osa_multi_*): one process, one host thread and
    stream per GPU, trajectories sharded by global id, one NCCL all-gather of the best records."""

    def __init__(self, handle, n):
        self._h = handle
        self.n = n
        self.nw = (n + 31) // 32
        c = ctypes.c_int()
        capi.check(capi.load().osa_multi_devices(self._h, ctypes.byref(c), None, 0))
        self.num_devices = c.value

    @staticmethod
    def _devices(devices, num_devices):
        if devices is None:
            return None, int(num_devices or 0)
        d = np.ascontiguousarray(devices, dtype=np.int32)
        return d, int(d.shape[0])

    @classmethod
    def dense(cls, qsym, devices=None, num_devices=0, sweep_precision=capi.SWEEP_F64):
        lib = capi.load()
        q = np.ascontiguousarray(qsym)
        if q.ndim != 2 or q.shape[0] != q.shape[1]:
            raise ValueError("qsym must be a square matrix")
        d, nd = cls._devices(devices, num_devices)
        h = ctypes.c_void_p()
        if q.dtype == np.float32:
            capi.check(lib.osa_multi_create_dense_f32(q.ctypes.data, q.shape[0],
                                                      d.ctypes.data if d is not None else None, nd,
                                                      ctypes.byref(h)))
        else:
            q = np.ascontiguousarray(q, dtype=np.float64)
            capi.check(lib.osa_multi_create_dense_f64(q.ctypes.data, q.shape[0],
                                                      d.ctypes.data if d is not None else None, nd,
                                                      sweep_precision, ctypes.byref(h)))
        return cls(h, q.shape[0])

    @classmethod
    def csr(cls, rowptr, col, val, diag, devices=None, num_devices=0,
            sweep_precision=capi.SWEEP_F64):
        lib = capi.load()
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        diag = np.ascontiguousarray(diag, dtype=np.float64)
        n = diag.shape[0]
        _check_csr_lengths(rowptr, col, val, n)
        d, nd = cls._devices(devices, num_devices)
        h = ctypes.c_void_p()
        capi.check(lib.osa_multi_create_csr_f64(rowptr.ctypes.data, col.ctypes.data, val.ctypes.data,
                                                diag.ctypes.data, n,
                                                d.ctypes.data if d is not None else None, nd,
                                                sweep_precision, ctypes.byref(h)))
        return cls(h, n)

    def close(self):
        if self._h:
            capi.load().osa_multi_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def anneal(self, beta_schedule, num_iter, num_tries, sweeps_per_beta=1, seed=1234,
               first_try=0, mode=capi.MODE_RANDOM_SITE, accept_rule=capi.ACCEPT_REFERENCE,
               kernel_variant=capi.KID_AUTO, want_energies=False, want_states=False):
        lib = capi.load()
        sched = np.ascontiguousarray(beta_schedule, dtype=np.float64)
        if sched.shape[0] < num_iter:
            raise ValueError("beta_schedule shorter than num_iter")
        prm = capi.AnnealParams(seed=seed, first_try=first_try, num_tries=num_tries,
                                num_iter=num_iter, sweeps_per_beta=sweeps_per_beta, mode=mode,
                                accept_rule=accept_rule, kernel_variant=kernel_variant, flags=0)
        energies = np.empty(num_tries, dtype=np.float64) if want_energies else None
        states = np.empty((num_tries, self.nw), dtype=np.uint32) if want_states else None
        state = np.empty(self.n, dtype=np.uint8)
        e = ctypes.c_double()
        idx = ctypes.c_uint64()
        st = capi.Stats()
        per = (capi.Stats * self.num_devices)()
        capi.check(lib.osa_multi_anneal(self._h, sched.ctypes.data, ctypes.byref(prm),
                                        energies.ctypes.data if want_energies else None,
                                        states.ctypes.data if want_states else None,
                                        state.ctypes.data, ctypes.byref(e), ctypes.byref(idx),
                                        ctypes.byref(st), ctypes.cast(per, ctypes.c_void_p)))
        res = AnnealResult(state, e.value, idx.value, st.as_dict(), energies, states)
        res.device_stats = [s.as_dict() for s in per]
        return res


def pinned_copy(array):
    """Copy `array` into page-locked host memory (osa_host_alloc_pinned); returns (ndarray, free)."""
    a = np.ascontiguousarray(array)
    ptr = ctypes.c_void_p()
    capi.check(capi.load().osa_host_alloc_pinned(a.nbytes, ctypes.byref(ptr)))
    buf = (ctypes.c_byte * a.nbytes).from_address(ptr.value)
    out = np.frombuffer(buf, dtype=a.dtype).reshape(a.shape)
    out[...] = a
    return out, (lambda: capi.load().osa_host_free_pinned(ptr))


def exhaustive(qsym, device=0):
    """exhaustive::solve on the GPU: (state uint8[N], energy) of the lowest-index ground state."""
    q = np.ascontiguousarray(qsym, dtype=np.float64)
    n = q.shape[0]
    state = np.empty(n, dtype=np.uint8)
    e = ctypes.c_double()
    capi.check(capi.load().osa_exhaustive_dense_f64(q.ctypes.data, n, device, state.ctypes.data,
                                                    ctypes.byref(e)))
    return state, e.value


def device_count():
    c = ctypes.c_int()
    capi.check(capi.load().osa_device_count(ctypes.byref(c)))
    return c.value


def device_name(device=0):
    buf = ctypes.create_string_buffer(256)
    capi.check(capi.load().osa_device_name(device, buf, 256))
    return buf.value.decode()


def measure_read_bandwidth(nbytes, iters=20, device=0):
    g = ctypes.c_double()
    capi.check(capi.load().osa_measure_read_bandwidth(device, nbytes, iters, ctypes.byref(g)))
    return g.value


def pack_states(states01):
    """[T][N] 0/1 -> [T][ceil(N/32)] uint32, bit i%32 of word i/32 = variable i."""
    s = np.asarray(states01, dtype=np.uint8)
    if s.ndim == 1:
        s = s.reshape(1, -1)
    t, n = s.shape
    nw = (n + 31) // 32
    padded = np.zeros((t, nw * 32), dtype=np.uint8)
    padded[:, :n] = s
    bits = padded.reshape(t, nw, 32).astype(np.uint64)
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64))
    return (bits * weights).sum(axis=2).astype(np.uint32)


def unpack_states(packed, n):
    p = np.asarray(packed, dtype=np.uint32)
    if p.ndim == 1:
        p = p.reshape(1, -1)
    bits = (p[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & np.uint32(1)
    return bits.reshape(p.shape[0], -1)[:, :n].astype(np.uint8)
