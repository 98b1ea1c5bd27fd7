"""The C-ABI library: loads without a GPU, exports every entry point include/taxila_gpu.h declares, and
fails loudly (no CPU fallback) when asked to compute without a CUDA device."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "taxila_gpu.h").read_text()
    return sorted(set(re.findall(r"TXG_API\s+[\w\s\*]+?\b(txg_\w+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    syms = declared_symbols()
    assert len(syms) == 28, syms
    for must in ("txg_create", "txg_set_walls", "txg_set_rho_u", "txg_set_fi", "txg_fi_init", "txg_update_moments", "txg_step",
                 "txg_collision", "txg_communicate_fi", "txg_stream", "txg_bounceback", "txg_apply_bcs", "txg_update_flux",
                 "txg_get_fi", "txg_get_state", "txg_get_diagnostics", "txg_delta_norm", "txg_destroy", "txg_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from taxila_lbm_b200 import capi

    lib = ctypes.CDLL(str(capi.LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes mirror binds exactly the header's surface
    assert sorted(capi.PROTOTYPES) == declared_symbols()
    capi.load()


def test_no_cpu_fallback():
    """Without a CUDA device txg_create must fail with a message, not compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import cases
    import taxila_lbm_b200 as tx

    cfg, _, _ = cases.bubble_2d(16, hw=3)
    with pytest.raises(Exception) as e:
        tx.Flow(cfg, device=0)
    assert "CUDA" in str(e.value) or "device" in str(e.value)


def test_product_never_imports_the_oracle():
    for f in (ROOT / "taxila-lbm_b200").glob("*.py"):
        assert "oracle" not in f.read_text().lower(), f
    for f in (ROOT / "taxila-lbm_b200" / "csrc").glob("*.cu*"):
        assert "taxila_oracle" not in f.read_text(), f
