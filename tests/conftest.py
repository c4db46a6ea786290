import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _have_gpu():
    try:
        from onesolver_b200 import capi
        import ctypes
        lib = capi.load()
        c = ctypes.c_int()
        return lib.osa_device_count(ctypes.byref(c)) == 0 and c.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (CPU checker) and make sure the CUDA library exists."""
    from oracle import binding
    binding.load()
    from onesolver_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def gpu():
    if not _have_gpu():
        pytest.fail("no CUDA device / library: GPU tests must run on the B200 box "
                    "(there is no CPU fallback)")
    return 0
