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
    assert lib.osa_abi_version() == 3


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


def test_multi_device_entry_has_no_cpu_fallback_either():
    lib = capi.load()
    c = ctypes.c_int(-1)
    if lib.osa_device_count(ctypes.byref(c)) == 0 and c.value > 0:
        pytest.skip("a CUDA device is present")
    q = np.zeros((4, 4))
    h = ctypes.c_void_p()
    rc = lib.osa_multi_create_dense_f64(q.ctypes.data, 4, None, 0, capi.SWEEP_F64, ctypes.byref(h))
    assert rc in (capi.OSA_ERR_NO_DEVICE, capi.OSA_ERR_CUDA) and not h.value
    assert lib.osa_multi_anneal(None, None, None, None, None, None, None, None, None, None) == capi.OSA_ERR_INVALID
    assert lib.osa_multi_destroy(None) == capi.OSA_OK


def test_csr_symmetry_is_checked_before_any_device_work():
    """An upper-triangle-only or value-asymmetric CSR is rejected (the dense path rejects an
    asymmetric Q the same way): the sweep and the exact energies would disagree on it."""
    lib = capi.load()
    h = ctypes.c_void_p()
    diag = np.zeros(3)
    # one-sided: (0,1) without (1,0)
    rowptr = np.array([0, 1, 1, 1], dtype=np.int32)
    col = np.array([1], dtype=np.int32)
    val = np.array([2.0])
    rc = lib.osa_problem_create_csr_f64(rowptr.ctypes.data, col.ctypes.data, val.ctypes.data,
                                        diag.ctypes.data, 3, 0, capi.SWEEP_F64, ctypes.byref(h))
    assert rc == capi.OSA_ERR_INVALID and b"not symmetric" in lib.osa_last_error()
    # both directions, different values
    rowptr = np.array([0, 1, 2, 2], dtype=np.int32)
    col = np.array([1, 0], dtype=np.int32)
    val = np.array([2.0, 3.0])
    rc = lib.osa_problem_create_csr_f64(rowptr.ctypes.data, col.ctypes.data, val.ctypes.data,
                                        diag.ctypes.data, 3, 0, capi.SWEEP_F64, ctypes.byref(h))
    assert rc == capi.OSA_ERR_INVALID and b"not symmetric" in lib.osa_last_error()


def test_python_wrapper_checks_csr_array_lengths():
    from onesolver_b200 import Problem
    with pytest.raises(ValueError):
        Problem.csr(np.array([0, 1], dtype=np.int32), np.array([1], dtype=np.int32), np.array([1.0]),
                    np.zeros(3))


def test_headline_kernel_keeps_its_accept_masks_on_the_uniform_datapath():
    """The dense sweep kernels owe a factor of 3.7 to ptxas keeping the accept masks -- and the
    +-1/0 multipliers derived from them -- in uniform registers (profiles/r02/
    ab_one_poller_uniform_datapath_lost.txt: innocuous-looking changes elsewhere in the kernel make
    it drop them, with bit-identical results).  Pin it on the built library: every packed FMA of the
    N = 4096 fp32 instantiations takes its multiplier from a uniform register."""
    import re
    import shutil
    import subprocess
    from onesolver_b200 import capi
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", capi.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    for kernel, expected in (("k_dense_seq_flowIfLi4ELi12ELi12ELi2ELb0", 192),
                             ("k_dense_seq_wsIfLi4ELi12ELi12ELi2ELb0ELi4E", 384)):
        m = re.search(r"Function : \S*" + kernel + r".*?(?=Function : |\Z)", out, re.S)
        assert m, f"{kernel} not found in the library"
        fma = re.findall(r"FFMA2 [^;]*;", m.group(0))
        uniform = [f for f in fma if re.search(r"UR\d+\.F32", f)]
        assert len(fma) == expected and len(uniform) == len(fma), (kernel, len(fma), len(uniform))
    # the warp-per-trajectory kernel has its three forms in both precisions: short rows, pipelined
    # flip-by-flip row add, rows of a batch added in one pass (osa_dense_generic.cu)
    for t in "fd":
        for form in ("Lb0ELb0E", "Lb1ELb0E", "Lb1ELb1E"):
            assert re.search(r"Function : \S*k_dense_genericI" + t + form, out), (t, form)
