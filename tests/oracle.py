"""ctypes wrapper of the CPU oracle (oracle/taxila_oracle.c).  TEST INFRASTRUCTURE:
imported only from tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB = ORACLE_DIR / "libtaxila_oracle.so"

import taxila_lbm_b200  # noqa: E402  (config mirror only)
from taxila_lbm_b200.config import TxgConfig  # noqa: E402

_lib = None


def build(force=False):
    src_m = max(p.stat().st_mtime for p in [ORACLE_DIR / "taxila_oracle.c", ORACLE_DIR / "ff_stencil_tables.h",
                                              ROOT / "include" / "taxila_gpu.h"])
    if force or not LIB.exists() or LIB.stat().st_mtime < src_m:
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.txo_create.restype = C.c_void_p
        L.txo_create.argtypes = [C.POINTER(TxgConfig)]
        L.txo_delta_norm.restype = C.c_double
        dp = C.POINTER(C.c_double)
        for name, args in {
            "txo_destroy": [],
            "txo_set_threads": [C.c_int],
            "txo_set_walls": [dp],
            "txo_get_walls_rg": [dp],
            "txo_set_rho": [dp],
            "txo_set_fi": [dp],
            "txo_fi_init": [],
            "txo_update_moments": [],
            "txo_step": [C.c_int],
            "txo_phase_collision": [],
            "txo_phase_communicate_fi": [],
            "txo_phase_stream": [],
            "txo_phase_bounceback": [],
            "txo_phase_apply_bcs": [],
            "txo_phase_update_flux": [],
            "txo_get_fi": [dp],
            "txo_get_rho": [dp],
            "txo_get_u": [dp],
            "txo_get_forces": [dp],
            "txo_diagnostics": [dp, dp, dp],
            "txo_delta_norm": [],
            "txo_set_bc_values": [C.c_int, dp],
            "txo_set_prestream": [C.c_int],
            "txo_set_bc_pressure_outlet": [C.c_int, C.c_double],
        }.items():
            fn = getattr(L, name)
            fn.argtypes = [C.c_void_p] + args
            if name != "txo_delta_norm":
                fn.restype = None
        L.txo_eos_bad.argtypes = [C.c_void_p]
        L.txo_eos_bad.restype = C.c_int
        L.txo_bc_supported.argtypes = [C.c_void_p]
        L.txo_bc_supported.restype = C.c_int
        ip = C.POINTER(C.c_int)
        L.txo_reflecting_pairs.argtypes = [C.c_void_p, C.c_int, ip, ip]
        L.txo_reflecting_pairs.restype = C.c_int
        L.txo_get_lattice.argtypes = [C.c_void_p, C.POINTER(C.c_int), dp, C.POINTER(C.c_int), dp, dp, dp]
        L.txo_get_lattice.restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """One single-domain CPU simulation in the reference's structure.
    Arrays cross in PETSc natural ordering: [z][y][x][...dofs], dofs with component fastest."""

    def __init__(self, cfg, threads=1):
        self.cfg = cfg.copy()
        self.cfg.zs, self.cfg.zl = 0, cfg.NZ
        self.L = lib()
        self.h = self.L.txo_create(C.byref(self.cfg))
        if not self.h:
            raise RuntimeError("txo_create failed (struct size mismatch?)")
        self.S, self.Q, self.D = cfg.ncomponents, cfg.Q, cfg.ndims
        self.NZ = cfg.NZ if cfg.ndims == 3 else 1
        self.NY, self.NX = cfg.NY, cfg.NX
        self.R = cfg.stencil_size_rho
        self.L.txo_set_threads(self.h, threads)

    def close(self):
        if self.h:
            self.L.txo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nodes(self):
        return (self.NZ, self.NY, self.NX)

    def set_walls(self, walls):
        w = np.ascontiguousarray(walls, dtype=np.float64).reshape(self.nodes)
        self.L.txo_set_walls(self.h, _dp(w))

    def walls_rg(self):
        R = self.R
        rz = R if self.D == 3 else 0
        out = np.empty((self.NZ + 2 * rz, self.NY + 2 * R, self.NX + 2 * R))
        self.L.txo_get_walls_rg(self.h, _dp(out))
        return out

    def shape_bc(self, boundary):
        n = {0: (self.NZ, self.NY), 1: (self.NZ, self.NX), 2: (self.NY, self.NX)}[boundary // 2]
        if self.D == 2:
            n = n[1:]
        return n + (self.D, self.S)

    def set_bc_values(self, boundary, vals):
        v = np.ascontiguousarray(vals, dtype=np.float64).reshape(self.shape_bc(boundary))
        self.L.txo_set_bc_values(self.h, int(boundary), _dp(v))

    def eos_bad(self):
        return self.L.txo_eos_bad(self.h)

    def set_bc_pressure_outlet(self, boundary, pressure):
        self.L.txo_set_bc_pressure_outlet(self.h, int(boundary), float(pressure))

    def set_prestream(self, on):
        self.L.txo_set_prestream(self.h, int(on))

    def reflecting_pairs(self, boundary):
        n = np.zeros(64, dtype=np.int32)
        p = np.zeros(64, dtype=np.int32)
        ip = C.POINTER(C.c_int)
        k = self.L.txo_reflecting_pairs(self.h, int(boundary), n.ctypes.data_as(ip), p.ctypes.data_as(ip))
        return list(zip(n[:k].tolist(), p[:k].tolist()))

    def set_rho(self, rho):
        r = np.ascontiguousarray(rho, dtype=np.float64).reshape(self.nodes + (self.S,))
        self.L.txo_set_rho(self.h, _dp(r))

    def set_fi(self, fi):
        f = np.ascontiguousarray(fi, dtype=np.float64).reshape(self.nodes + (self.Q, self.S))
        self.L.txo_set_fi(self.h, _dp(f))

    def fi_init(self):
        self.L.txo_fi_init(self.h)

    def update_moments(self):
        self.L.txo_update_moments(self.h)

    def step(self, n=1):
        self.L.txo_step(self.h, int(n))

    def phase(self, name):
        getattr(self.L, "txo_phase_" + name)(self.h)

    def fi(self):
        out = np.empty(self.nodes + (self.Q, self.S))
        self.L.txo_get_fi(self.h, _dp(out))
        return out

    def rho(self):
        out = np.empty(self.nodes + (self.S,))
        self.L.txo_get_rho(self.h, _dp(out))
        return out

    def u(self):
        out = np.empty(self.nodes + (self.D, self.S))
        self.L.txo_get_u(self.h, _dp(out))
        return out

    def forces(self):
        out = np.empty(self.nodes + (self.D, self.S))
        self.L.txo_get_forces(self.h, _dp(out))
        return out

    def diagnostics(self):
        rhot = np.empty(self.nodes)
        prs = np.empty(self.nodes)
        velt = np.empty(self.nodes + (self.D,))
        self.L.txo_diagnostics(self.h, _dp(rhot), _dp(prs), _dp(velt))
        return rhot, prs, velt

    def delta_norm(self):
        return self.L.txo_delta_norm(self.h)

    def lattice(self):
        Q = self.Q
        ci = np.zeros((Q, 3), dtype=np.int32)
        w = np.zeros(Q)
        opp = np.zeros(Q, dtype=np.int32)
        mt = np.zeros((Q, Q))
        mmt = np.zeros(Q)
        ffw = np.zeros(41)
        self.L.txo_get_lattice(self.h, ci.ctypes.data_as(C.POINTER(C.c_int)), _dp(w),
                               opp.ctypes.data_as(C.POINTER(C.c_int)), _dp(mt), _dp(mmt), _dp(ffw))
        return dict(ci=ci, weights=w, opposites=opp, mt=mt, mmt=mmt, ffw=ffw)
