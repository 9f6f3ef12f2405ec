"""libfluid_b200 -- B200-native implementation of libfluid's per-step simulation hot path.

Layout:  csrc/ (CUDA kernels + the C ABI of include/lfk.h)  ·  capi.py (ctypes handle on that ABI)  ·
build.py (nvcc, sm_100a, in-tree)  ·  host/ + include/fluid/ (the C++ mirror of fluid::simulation).
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
