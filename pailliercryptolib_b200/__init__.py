"""B200-native (sm_100a) back-end for the modexp hot path of
intel/pailliercryptolib: hand-written CUDA kernels behind a C ABI
(include/ipcl_b200.h) plus the C++ `ipcl::` host API mirror.

Python here is only the build driver and the ctypes harness the tests and
bench.py use; the product is the two shared libraries under lib/."""
from . import build  # noqa: F401

__all__ = ["build", "capi", "limbs"]
