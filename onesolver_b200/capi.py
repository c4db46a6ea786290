"""ctypes binding of the C ABI in include/onesolver_b200.h (test/bench driver only).

The product is the shared library; this module is the thinnest possible way for
pytest and bench.py to call it.  It never computes anything itself and raises
if the CUDA library is missing -- there is no CPU fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OSA_LIB_PATH: A/B runs of tools/probe.py against another build of the same library
LIB_PATH = os.environ.get("OSA_LIB_PATH") or os.path.join(_HERE, "lib", "libonesolver_b200.so")

OSA_OK, OSA_ERR_INVALID, OSA_ERR_CUDA, OSA_ERR_NO_DEVICE, OSA_ERR_UNSUPPORTED, OSA_ERR_NOMEM = range(6)
MODE_RANDOM_SITE, MODE_SEQUENTIAL_SWEEP = 0, 1
ACCEPT_REFERENCE, ACCEPT_BOLTZMANN = 0, 1
SWEEP_F64, SWEEP_F32 = 0, 1
KID_AUTO, KID_DENSE_SEQ, KID_DENSE_GENERIC, KID_SPARSE = 0, 1, 2, 3

# every symbol include/onesolver_b200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "osa_abi_version", "osa_last_error", "osa_device_count", "osa_device_name",
    "osa_kernel_name", "osa_problem_create_dense_f64", "osa_problem_create_dense_f32",
    "osa_problem_create_csr_f64", "osa_problem_destroy", "osa_problem_size", "osa_anneal",
    "osa_anneal_traced",
    "osa_pt_anneal", "osa_pa_anneal", "osa_energy_batch", "osa_exhaustive_dense_f64", "osa_host_alloc_pinned",
    "osa_host_free_pinned", "osa_measure_read_bandwidth",
    "osa_multi_create_dense_f64", "osa_multi_create_dense_f32", "osa_multi_create_csr_f64",
    "osa_multi_destroy", "osa_multi_devices", "osa_multi_problem", "osa_multi_anneal",
]


class AnnealParams(ctypes.Structure):
    _fields_ = [
        ("seed", ctypes.c_uint64),
        ("first_try", ctypes.c_uint64),
        ("num_tries", ctypes.c_uint64),
        ("num_iter", ctypes.c_int32),
        ("sweeps_per_beta", ctypes.c_int32),
        ("mode", ctypes.c_int32),
        ("accept_rule", ctypes.c_int32),
        ("kernel_variant", ctypes.c_int32),
        ("flags", ctypes.c_int32),
    ]


class PtParams(ctypes.Structure):
    _fields_ = [
        ("seed", ctypes.c_uint64),
        ("first_group", ctypes.c_uint64),
        ("num_groups", ctypes.c_uint64),
        ("num_replicas", ctypes.c_int32),
        ("num_rounds", ctypes.c_int32),
        ("sweeps_per_round", ctypes.c_int32),
        ("accept_rule", ctypes.c_int32),
        ("flags", ctypes.c_uint32),
        ("reserved", ctypes.c_int32),
    ]


class PaParams(ctypes.Structure):
    _fields_ = [
        ("seed", ctypes.c_uint64),
        ("first_population", ctypes.c_uint64),
        ("num_populations", ctypes.c_uint64),
        ("population_size", ctypes.c_int32),
        ("num_steps", ctypes.c_int32),
        ("sweeps_per_step", ctypes.c_int32),
        ("accept_rule", ctypes.c_int32),
        ("flags", ctypes.c_uint32),
        ("reserved", ctypes.c_int32),
    ]


class Stats(ctypes.Structure):
    _fields_ = [
        ("attempts", ctypes.c_uint64),
        ("accepts", ctypes.c_uint64),
        ("row_fetches", ctypes.c_uint64),
        ("init_row_fetches", ctypes.c_uint64),
        ("ms_total", ctypes.c_float),
        ("ms_sweep", ctypes.c_float),
        ("ms_energy", ctypes.c_float),
        ("ms_reduce", ctypes.c_float),
        ("kernel_id", ctypes.c_int32),
        ("traj_per_batch", ctypes.c_int32),
        ("q_elem_bytes", ctypes.c_int32),
        ("grid", ctypes.c_int32),
        ("launches", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("cyc_decide", ctypes.c_uint64),
        ("cyc_apply", ctypes.c_uint64),
        ("cyc_stage", ctypes.c_uint64),
        ("cyc_init", ctypes.c_uint64),
        ("pt_swaps", ctypes.c_uint64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class OsaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"osa error {code}: {message}")
        self.code = code


_lib = None


def _preload_bundled_nccl():
    """The library binds NCCL at run time by soname (osa_multi.cu), which returns the copy the
    process already has.  A Python host may import torch LATER, and torch needs the NCCL it was
    built with (its wheel's nvidia/nccl/lib/libnccl.so.2; the system copy is older and lacks
    symbols) -- whichever libnccl.so.2 is loaded first wins.  So a Python process loads the bundled
    copy first, when there is one; torch is not imported for it.  C/C++ hosts without torch simply
    get the system NCCL."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations if spec else []):
            path = os.path.join(loc, "lib", "libnccl.so.2")
            if os.path.exists(path):
                ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
                return path
    except Exception:  # a missing or unloadable bundled copy is not an error here
        pass
    return None


def load():
    """Load libonesolver_b200.so; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OsaError(-1, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (no CPU fallback exists)")
    _preload_bundled_nccl()
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, u64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_size_t
    P = ctypes.POINTER
    lib.osa_abi_version.restype = i32
    lib.osa_last_error.restype = ctypes.c_char_p
    lib.osa_device_count.argtypes = [P(i32)]
    lib.osa_device_name.argtypes = [i32, ctypes.c_char_p, sz]
    lib.osa_kernel_name.argtypes = [i32]
    lib.osa_kernel_name.restype = ctypes.c_char_p
    lib.osa_problem_create_dense_f64.argtypes = [vp, i32, i32, i32, P(vp)]
    lib.osa_problem_create_dense_f32.argtypes = [vp, i32, i32, P(vp)]
    lib.osa_problem_create_csr_f64.argtypes = [vp, vp, vp, vp, i32, i32, i32, P(vp)]
    lib.osa_problem_destroy.argtypes = [vp]
    lib.osa_problem_size.argtypes = [vp, P(i32), P(i32), P(i32)]
    lib.osa_anneal.argtypes = [vp, vp, P(AnnealParams), vp, vp, vp, P(ctypes.c_double), P(u64),
                               P(Stats)]
    lib.osa_anneal_traced.argtypes = [vp, vp, P(AnnealParams), vp, vp, vp, P(ctypes.c_double), P(u64),
                                      vp, P(Stats)]
    lib.osa_pt_anneal.argtypes = [vp, vp, P(PtParams), vp, vp, vp, P(ctypes.c_double), P(u64),
                                  P(Stats)]
    lib.osa_pa_anneal.argtypes = [vp, vp, P(PaParams), vp, vp, vp, P(ctypes.c_double), P(u64),
                                  P(Stats)]
    lib.osa_energy_batch.argtypes = [vp, vp, u64, vp]
    lib.osa_host_alloc_pinned.argtypes = [sz, P(vp)]
    lib.osa_host_free_pinned.argtypes = [vp]
    lib.osa_exhaustive_dense_f64.argtypes = [vp, i32, i32, vp, P(ctypes.c_double)]
    lib.osa_measure_read_bandwidth.argtypes = [i32, sz, i32, P(ctypes.c_double)]
    lib.osa_multi_create_dense_f64.argtypes = [vp, i32, vp, i32, i32, P(vp)]
    lib.osa_multi_create_dense_f32.argtypes = [vp, i32, vp, i32, P(vp)]
    lib.osa_multi_create_csr_f64.argtypes = [vp, vp, vp, vp, i32, vp, i32, i32, P(vp)]
    lib.osa_multi_destroy.argtypes = [vp]
    lib.osa_multi_devices.argtypes = [vp, P(i32), vp, i32]
    lib.osa_multi_problem.argtypes = [vp, i32, P(vp)]
    lib.osa_multi_anneal.argtypes = [vp, vp, P(AnnealParams), vp, vp, vp, P(ctypes.c_double), P(u64),
                                     P(Stats), vp]
    _lib = lib
    return lib


def check(rc):
    if rc != OSA_OK:
        raise OsaError(rc, load().osa_last_error().decode("utf-8", "replace"))
