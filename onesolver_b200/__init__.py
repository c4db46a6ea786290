"""onesolver_b200 -- B200-native simulated-annealing engine behind oneSolver's API.

The product is `lib/libonesolver_b200.so` (C ABI in include/onesolver_b200.h, kernels in
csrc/).  The Python modules here only drive it from pytest and bench.py.
"""
from . import capi  # noqa: F401
from .anneal import (  # noqa: F401
    AnnealResult,
    MultiProblem,
    Problem,
    construct_geometric_beta_schedule,
    construct_linear_beta_schedule,
    device_count,
    device_name,
    exhaustive,
    measure_read_bandwidth,
    pack_states,
    pinned_copy,
    unpack_states,
)
