"""gaussianvi_b200 -- B200-native (sm_100a, FP64) factorized natural-gradient GVI hot path.

The product is the CUDA shared library libgvib200.so (C-ABI: include/gvib200.h) plus the C++ facade
under gaussianvi_b200/cpp/gvi that keeps the reference's class names.  This Python package is the
ctypes mirror used by tests and bench.py; it contains no numerical fallback."""
from . import capi  # noqa: F401
from .capi import Context, Problem, GviError, load_library  # noqa: F401
