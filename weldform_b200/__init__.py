"""weldform_b200 -- B200-native implementation of WeldForm's per-timestep SPH hot path.

The product is libweldform_b200.so (hand-written sm_100a CUDA kernels behind the C ABI in include/weldform_b200.h).
This package is the thin Python binding (ctypes) used by the tests and bench.py; the C++ mirror of the reference's
Domain/Particle source API lives in weldform_b200/host/. There is no CPU fallback: importing works without a GPU (so
that the build can be checked), but creating a Context without the library or without a device raises.
"""
from .api import Context, WeldformError, lib_path, load_library, FIELDS  # noqa: F401

__all__ = ["Context", "WeldformError", "lib_path", "load_library", "FIELDS"]
