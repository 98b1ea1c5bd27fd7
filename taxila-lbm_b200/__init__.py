"""taxila-lbm_b200: the B200-native flow hot path of Taxila-LBM behind a C ABI.

Python here is host-side plumbing only (configuration, synthetic inputs, the
ctypes binding of libtaxila_gpu.so); the product is the CUDA library in csrc/.
"""
from . import config  # noqa: F401
