"""ctypes binding of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import this module (see oracle/osa_oracle.h).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

MODE_RANDOM_SITE, MODE_SEQUENTIAL_SWEEP = 0, 1
STREAM_INIT, STREAM_SEQ, STREAM_RND = 0, 1, 2


class Counters(ctypes.Structure):
    _fields_ = [("attempts", ctypes.c_uint64), ("accepts", ctypes.c_uint64),
                ("row_fetches", ctypes.c_uint64), ("init_row_fetches", ctypes.c_uint64)]


_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, u32, u64, i64, dbl, sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32,
                                       ctypes.c_uint64, ctypes.c_int64, ctypes.c_double,
                                       ctypes.c_size_t)
    P = ctypes.POINTER
    lib.orc_philox4x32_10.argtypes = [vp, vp, vp]
    lib.orc_engine_draw.argtypes = [u64, u64, u32, u32, u32, vp]
    lib.orc_set_trace_output.argtypes = [vp]
    lib.orc_set_trace_output.restype = None
    lib.orc_det_exp.argtypes = [ctypes.c_double]
    lib.orc_det_exp.restype = ctypes.c_double
    lib.orc_pa_resample.argtypes = [vp, ctypes.c_int, ctypes.c_double, u64, u64, u32, vp]
    lib.orc_pa_resample.restype = None
    lib.orc_neglogf.argtypes = [u32]
    lib.orc_neglogf.restype = ctypes.c_float
    lib.orc_init_bit.argtypes = [u64, u64, u32]
    lib.orc_init_bit.restype = i32
    lib.orc_ref_energy.argtypes = [vp, vp, i32]
    lib.orc_ref_energy.restype = dbl
    lib.orc_ref_flatten.argtypes = [i32, vp, vp, i32, vp, vp, vp, i32, vp]
    lib.orc_ref_schedule_linear.argtypes = [vp, dbl, dbl, ctypes.c_uint]
    lib.orc_ref_schedule_geometric.argtypes = [vp, dbl, dbl, ctypes.c_uint]
    lib.orc_ref_anneal.argtypes = [vp, i32, vp, i32, u64, i32, u64, u64, vp, vp, i32]
    lib.orc_ref_anneal.restype = i64
    lib.orc_ref_exhaustive.argtypes = [vp, i32, i32, vp, P(dbl)]
    lib.orc_ref_solution_csv.argtypes = [vp, i32, dbl, ctypes.c_char_p, sz]
    lib.orc_ref_solution_csv.restype = sz
    dense = [vp, vp, i32, sz, vp, i32, i32, i32, u64, u64, u64, i32, vp, vp, vp, P(Counters)]
    lib.orc_replay_dense_f32.argtypes = dense
    lib.orc_replay_dense_f64.argtypes = dense
    lib.orc_replay_dense_f32_round.argtypes = dense + [vp, vp, u32]
    lib.orc_replay_dense_f64_round.argtypes = dense + [vp, vp, u32]
    csr = [vp, vp, vp, vp, i32, vp, i32, i32, i32, u64, u64, u64, vp, vp, vp, P(Counters)]
    lib.orc_replay_csr_f32.argtypes = csr
    lib.orc_replay_csr_f64.argtypes = csr
    lib.orc_energy_packed.argtypes = [vp, i32, vp, u64, vp]
    lib.orc_num_threads.restype = i32
    _lib = lib
    return lib


# ---- thin numpy wrappers ---------------------------------------------------

def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    load().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def neglogf(w):
    return float(load().orc_neglogf(int(w)))


def det_exp(x):
    return load().orc_det_exp(float(x))


def pa_resample(energies, neg_db, seed, population, step):
    """Source replica of every slot after one resampling step of osa_pa_anneal."""
    e = np.ascontiguousarray(energies, dtype=np.float64)
    src = np.zeros(e.shape[0], dtype=np.int32)
    load().orc_pa_resample(e.ctypes.data, e.shape[0], float(neg_db), seed, population, step,
                           src.ctypes.data)
    return src


def ref_energy(flat_qubo, state):
    q = np.ascontiguousarray(flat_qubo, dtype=np.float64)
    s = np.ascontiguousarray(state, dtype=np.int8)
    return load().orc_ref_energy(q.ctypes.data, s.ctypes.data, s.shape[0])


def ref_flatten(n, linear, quadratic):
    """linear: {i: v}; quadratic: {(i, j): v} -> n*n float64 (flatten_qubo layout)."""
    li = np.array(list(linear.keys()), dtype=np.int32)
    lv = np.array(list(linear.values()), dtype=np.float64)
    qi = np.array([k[0] for k in quadratic.keys()], dtype=np.int32)
    qj = np.array([k[1] for k in quadratic.keys()], dtype=np.int32)
    qv = np.array(list(quadratic.values()), dtype=np.float64)
    out = np.zeros(n * n, dtype=np.float64)
    load().orc_ref_flatten(n, li.ctypes.data, lv.ctypes.data, len(li), qi.ctypes.data,
                           qj.ctypes.data, qv.ctypes.data, len(qi), out.ctypes.data)
    return out


def ref_schedule(kind, beta_min, beta_max, num_iter):
    out = np.zeros(num_iter, dtype=np.float64)
    fn = load().orc_ref_schedule_linear if kind == "linear" else load().orc_ref_schedule_geometric
    fn(out.ctypes.data, beta_min, beta_max, num_iter)
    return out


def ref_anneal(flat_qubo, n, schedule, num_iter, num_tries, sweeps_per_beta=1, seed=1234,
               first_try=0, num_threads=0):
    q = np.ascontiguousarray(flat_qubo, dtype=np.float64)
    s = np.ascontiguousarray(schedule, dtype=np.float64)
    states = np.zeros((num_tries, n), dtype=np.int8)
    energies = np.zeros(num_tries, dtype=np.float64)
    idx = load().orc_ref_anneal(q.ctypes.data, n, s.ctypes.data, num_iter, num_tries,
                                sweeps_per_beta, seed, first_try, states.ctypes.data,
                                energies.ctypes.data, num_threads)
    return int(idx), states, energies


def ref_exhaustive(flat_qubo, n, num_ranges=8):
    q = np.ascontiguousarray(flat_qubo, dtype=np.float64)
    state = np.zeros(n, dtype=np.int8)
    e = ctypes.c_double()
    rc = load().orc_ref_exhaustive(q.ctypes.data, n, num_ranges, state.ctypes.data,
                                   ctypes.byref(e))
    if rc != 0:
        raise ValueError("exhaustive oracle: n must be in [1, 30]")
    return state, e.value


def ref_solution_csv(state, energy):
    s = np.ascontiguousarray(state, dtype=np.int8)
    buf = ctypes.create_string_buffer(16 * (len(s) + 4) + 64)
    n = load().orc_ref_solution_csv(s.ctypes.data, len(s), energy, buf, len(buf))
    return buf.raw[:n].decode()


def split_dense(qsym, dtype, ld=None):
    """flatten_qubo layout -> (zero-diagonal matrix [n][ld], diag[n]) in `dtype`."""
    q = np.asarray(qsym, dtype=np.float64)
    n = q.shape[0]
    ld = ld or n
    qoff = np.zeros((n, ld), dtype=dtype)
    qoff[:, :n] = q.astype(dtype)
    qoff[np.arange(n), np.arange(n)] = 0
    diag = np.diag(q).astype(dtype).copy()
    return qoff, diag


def tscale(schedule, accept_rule, dtype):
    s = np.asarray(schedule, dtype=np.float64)
    t = s if accept_rule == 0 else 1.0 / s
    return np.ascontiguousarray(t.astype(dtype))


class _Trace:
    """with _Trace(n) as t: ... replay ...; t.hashes = flip traces of the n trajectories."""

    def __init__(self, num_tries):
        self.hashes = np.zeros(num_tries, dtype=np.uint64)

    def __enter__(self):
        load().orc_set_trace_output(ctypes.c_void_p(self.hashes.ctypes.data))
        return self

    def __exit__(self, *exc):
        load().orc_set_trace_output(None)


def trace(num_tries):
    return _Trace(num_tries)


def replay_dense(qsym, schedule, num_iter, num_tries, sweeps_per_beta=1, mode=0, accept_rule=0,
                 seed=1234, first_try=0, dtype=np.float64, batch_r=1, want_final=False):
    """Bit-exact host replay of the CUDA engine on a dense problem."""
    lib = load()
    qoff, diag = split_dense(qsym, dtype)
    n = qoff.shape[0]
    nw = (n + 31) // 32
    ts = tscale(schedule, accept_rule, dtype)
    best_rel = np.zeros(num_tries, dtype=np.float64)
    best = np.zeros((num_tries, nw), dtype=np.uint32)
    final = np.zeros((num_tries, nw), dtype=np.uint32) if want_final else None
    cnt = Counters()
    fn = lib.orc_replay_dense_f32 if dtype == np.float32 else lib.orc_replay_dense_f64
    rc = fn(qoff.ctypes.data, diag.ctypes.data, n, qoff.shape[1], ts.ctypes.data, num_iter,
            sweeps_per_beta, mode, seed, first_try, num_tries, batch_r, best_rel.ctypes.data,
            best.ctypes.data, final.ctypes.data if want_final else None, ctypes.byref(cnt))
    if rc != 0:
        raise ValueError("orc_replay_dense failed")
    return best_rel, best, final, cnt


def engine_draw(seed, traj, stream, c0, c1):
    out = np.zeros(4, dtype=np.uint32)
    load().orc_engine_draw(int(seed), int(traj), int(stream), int(c0), int(c1), out.ctypes.data)
    return out


def replay_dense_round(qoff, diag, ts_traj, init_states, sweeps, seed, first_try, step_base):
    """One resumable launch of the dense sweep kernel (sequential sweeps, per-trajectory
    threshold scale, start states given): -> (best_rel, best_states, final_states)."""
    lib = load()
    dtype = qoff.dtype.type
    n = qoff.shape[0]
    nw = (n + 31) // 32
    tries = init_states.shape[0]
    ts = np.ascontiguousarray(ts_traj, dtype=dtype)
    init = np.ascontiguousarray(init_states, dtype=np.uint32)
    best_rel = np.zeros(tries, dtype=np.float64)
    best = np.zeros((tries, nw), dtype=np.uint32)
    final = np.zeros((tries, nw), dtype=np.uint32)
    fn = lib.orc_replay_dense_f32_round if dtype == np.float32 else lib.orc_replay_dense_f64_round
    rc = fn(qoff.ctypes.data, diag.ctypes.data, n, qoff.shape[1], None, 1, sweeps, 1, seed,
            first_try, tries, 1, best_rel.ctypes.data, best.ctypes.data, final.ctypes.data, None,
            init.ctypes.data, ts.ctypes.data, step_base)
    if rc != 0:
        raise ValueError("orc_replay_dense_round failed")
    return best_rel, best, final


def replay_csr(rowptr, col, val, diag, schedule, num_iter, num_tries, sweeps_per_beta=1, mode=1,
               accept_rule=0, seed=1234, first_try=0, dtype=np.float64, want_final=False):
    lib = load()
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    v = np.ascontiguousarray(np.asarray(val, dtype=np.float64).astype(dtype))
    d = np.ascontiguousarray(np.asarray(diag, dtype=np.float64).astype(dtype))
    n = d.shape[0]
    nw = (n + 31) // 32
    ts = tscale(schedule, accept_rule, dtype)
    best_rel = np.zeros(num_tries, dtype=np.float64)
    best = np.zeros((num_tries, nw), dtype=np.uint32)
    final = np.zeros((num_tries, nw), dtype=np.uint32) if want_final else None
    cnt = Counters()
    fn = lib.orc_replay_csr_f32 if dtype == np.float32 else lib.orc_replay_csr_f64
    rc = fn(rowptr.ctypes.data, col.ctypes.data, v.ctypes.data, d.ctypes.data, n, ts.ctypes.data,
            num_iter, sweeps_per_beta, mode, seed, first_try, num_tries, best_rel.ctypes.data,
            best.ctypes.data, final.ctypes.data if want_final else None, ctypes.byref(cnt))
    if rc != 0:
        raise ValueError("orc_replay_csr failed")
    return best_rel, best, final, cnt


def energy_packed(qsym, states_packed):
    q = np.ascontiguousarray(qsym, dtype=np.float64)
    s = np.ascontiguousarray(states_packed, dtype=np.uint32)
    out = np.zeros(s.shape[0], dtype=np.float64)
    load().orc_energy_packed(q.ctypes.data, q.shape[0], s.ctypes.data, s.shape[0],
                             out.ctypes.data)
    return out


def num_threads():
    return load().orc_num_threads()
