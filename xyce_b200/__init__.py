"""xyce_b200 -- B200-native Newton-step engine behind Xyce's device / loader / linear-solver contracts.

The product is the CUDA library `xyce_b200/lib/libxyce_b200.so` with the C ABI of
`include/xyce_b200.h`.  This Python package is thin plumbing around that ABI (ctypes) used by the
tests and the benchmark; it contains no numerical fallback and raises if the library is missing.
"""
from .capi import Engine, SolverState, TranParams, load_library, LIB_PATH  # noqa: F401
