"""taxila-lbm_b200: the B200-native flow hot path of Taxila-LBM behind a C ABI.

Python here is host-side plumbing only (configuration, synthetic inputs, the ctypes binding of
libtaxila_gpu.so); the product is the CUDA library built from csrc/.
"""
from . import config, geometry, petsc_io, slab, workloads  # noqa: F401
from .flow import Flow  # noqa: F401
