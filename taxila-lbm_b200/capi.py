"""ctypes binding of libtaxila_gpu.so -- exactly the entry points include/taxila_gpu.h declares.

This is what a reference-side binding looks like from Python; the Fortran ISO_C_BINDING
equivalent is in shim/lbm_gpu_binding.F90 and INTEGRATION.md.  There is no fallback: if
the CUDA library is missing or a call fails, an exception carries txg_last_error().
"""
import ctypes as C
from pathlib import Path

from .config import TxgConfig

import os

# TAXILA_GPU_LIB selects another build of the same library (kernel-tuning experiments)
LIB_PATH = Path(os.environ.get("TAXILA_GPU_LIB") or Path(__file__).resolve().parent / "libtaxila_gpu.so")

_dp = C.POINTER(C.c_double)
_h = C.c_void_p

# name -> (restype, argtypes); mirrors include/taxila_gpu.h one to one
PROTOTYPES = {
    "txg_config_defaults": (C.c_int, [C.POINTER(TxgConfig)]),
    "txg_create": (C.c_int, [C.POINTER(_h), C.POINTER(TxgConfig), C.c_int]),
    "txg_destroy": (C.c_int, [_h]),
    "txg_last_error": (C.c_char_p, [_h]),
    "txg_nccl_unique_id": (C.c_int, [C.POINTER(C.c_ubyte)]),
    "txg_comm_init": (C.c_int, [_h, C.POINTER(C.c_ubyte)]),
    "txg_set_walls": (C.c_int, [_h, _dp]),
    "txg_set_bc_values": (C.c_int, [_h, C.c_int, _dp]),
    "txg_set_bc_pressure_outlet": (C.c_int, [_h, C.c_int, C.c_double]),
    "txg_set_rho_u": (C.c_int, [_h, _dp, _dp]),
    "txg_set_fi": (C.c_int, [_h, _dp]),
    "txg_fi_init": (C.c_int, [_h]),
    "txg_update_moments": (C.c_int, [_h]),
    "txg_step": (C.c_int, [_h, C.c_int]),
    "txg_collision": (C.c_int, [_h]),
    "txg_communicate_fi": (C.c_int, [_h]),
    "txg_stream": (C.c_int, [_h]),
    "txg_bounceback": (C.c_int, [_h]),
    "txg_apply_bcs": (C.c_int, [_h]),
    "txg_update_flux": (C.c_int, [_h]),
    "txg_get_fi": (C.c_int, [_h, _dp]),
    "txg_get_state": (C.c_int, [_h, _dp, _dp, _dp]),
    "txg_get_diagnostics": (C.c_int, [_h, _dp, _dp, _dp]),
    "txg_get_node_class": (C.c_int, [_h, C.POINTER(C.c_uint8)]),
    "txg_delta_norm": (C.c_int, [_h, _dp]),
    "txg_synchronize": (C.c_int, [_h]),
    "txg_last_step_ms": (C.c_int, [_h, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "txg_enable_kernel_timing": (C.c_int, [_h, C.c_int]),
    "txg_kernel_times": (C.c_int, [_h, C.c_int, C.POINTER(C.c_char_p), _dp, C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "txg_reset_kernel_times": (C.c_int, [_h]),
}

_lib = None


class TaxilaGpuError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("taxila_gpu error %d: %s" % (code, message))
        self.code = code


def load():
    """Load the CUDA library.  Raises if it has not been built -- there is no other path."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  taxila-lbm_b200 has no CPU or PyTorch fallback." % LIB_PATH
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(lib, handle, rc):
    if rc != 0:
        msg = lib.txg_last_error(handle)
        raise TaxilaGpuError(rc, msg.decode() if msg else "")
