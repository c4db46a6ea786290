"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares,
and refuses to compute without a CUDA device (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from onesolver_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "onesolver_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(osa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = header_symbols()
    assert len(names) >= 13
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == names
    assert lib.osa_abi_version() == 2


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(capi.AnnealParams) == 48
    assert ctypes.sizeof(capi.Stats) == 112
    assert ctypes.sizeof(capi.PtParams) == 48


def test_no_cpu_fallback_without_a_device():
    lib = capi.load()
    c = ctypes.c_int(-1)
    rc = lib.osa_device_count(ctypes.byref(c))
    if rc == 0 and c.value > 0:
        pytest.skip("a CUDA device is present")
    q = np.zeros((4, 4))
    h = ctypes.c_void_p()
    rc = lib.osa_problem_create_dense_f64(q.ctypes.data, 4, 0, capi.SWEEP_F64, ctypes.byref(h))
    assert rc in (capi.OSA_ERR_NO_DEVICE, capi.OSA_ERR_CUDA) and not h.value
    assert lib.osa_last_error()  # a message is always available
    g = ctypes.c_double()
    assert lib.osa_measure_read_bandwidth(0, 1 << 20, 1, ctypes.byref(g)) != 0


def test_argument_validation_needs_no_device():
    lib = capi.load()
    h = ctypes.c_void_p()
    assert lib.osa_problem_create_dense_f64(None, 4, 0, 0, ctypes.byref(h)) == capi.OSA_ERR_INVALID
    q = np.zeros((4, 4))
    assert lib.osa_problem_create_dense_f64(q.ctypes.data, 0, 0, 0, ctypes.byref(h)) == capi.OSA_ERR_INVALID
    assert lib.osa_problem_create_dense_f64(q.ctypes.data, 4, 0, 7, ctypes.byref(h)) == capi.OSA_ERR_INVALID
    assert lib.osa_anneal(None, None, None, None, None, None, None, None, None) == capi.OSA_ERR_INVALID
    assert lib.osa_kernel_name(1) == b"dense_seq"
