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
    assert len(syms) == 30, syms
    for must in ("txg_create", "txg_set_walls", "txg_set_bc_values", "txg_set_rho_u", "txg_set_fi", "txg_fi_init", "txg_update_moments", "txg_step",
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


def _c_config_fields():
    text = (ROOT / "include" / "taxila_gpu.h").read_text()
    body = text[text.index("typedef struct txg_config {") + len("typedef struct txg_config {"):text.index("} txg_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        m = re.match(r"\s*(?:int32_t|double)\s+(.*)$", decl.strip(), flags=re.S)
        if m:
            for part in m.group(1).split(","):
                names.append(re.match(r"\s*(\w+)", part).group(1))
    return names


def _f90_config_fields():
    text = (ROOT / "shim" / "lbm_gpu_binding.F90").read_text()
    body = text[text.index("type, bind(C), public :: txg_config"):text.index("end type txg_config")]
    names = []
    for line in body.splitlines()[1:]:
        line = line.split("!")[0]
        if "::" in line:
            for part in re.split(r",(?![^()]*\))", line.split("::")[1]):
                names.append(re.match(r"\s*(\w+)", part).group(1))
    return names


def test_fortran_binding_mirrors_the_header():
    """shim/lbm_gpu_binding.F90 (not compilable in this image): same struct fields in the same order, and one
    bind(C) interface per exported function."""
    assert _f90_config_fields() == _c_config_fields()
    text = (ROOT / "shim" / "lbm_gpu_binding.F90").read_text()
    bound = sorted(set(re.findall(r'bind\(C, name="(txg_\w+)"\)', text)))
    assert bound == declared_symbols()
