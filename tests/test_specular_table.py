"""Free-slip walls (WALL_NORMAL_X/Y/Z = 900-902) without a GPU: the host-side table builder of the CUDA
library (csrc/specular_table.h, compiled here with g++) + a numpy replay of what the device does with it
(push with plain bounce-back, then dst <- src) against the oracle's literal communicate / stream /
bounce-back sweep (lbm_distribution_function.F90:560-784), bit for bit."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import cases  # noqa: F401
import oracle
from taxila_lbm_b200 import config as tc
from taxila_lbm_b200 import geometry as geo

HERE = Path(__file__).resolve().parent
SRC = HERE / "native" / "specular_table_capi.cpp"
HDR = HERE.parent / "taxila-lbm_b200" / "csrc" / "specular_table.h"
LIB = HERE / "native" / "_build" / "libspecular_table.so"


def _lib():
    LIB.parent.mkdir(exist_ok=True)
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, HDR.stat().st_mtime):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(LIB), str(SRC)], check=True)
    L = C.CDLL(str(LIB))
    ip, up = C.POINTER(C.c_int), C.POINTER(C.c_uint32)
    L.spec_build.restype = C.c_longlong
    L.spec_build.argtypes = [C.c_int, C.c_int, ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, C.POINTER(C.c_uint8), up,
                             C.c_longlong, up, up, C.c_longlong, C.POINTER(C.c_longlong)]
    return L


def _classes(walls_rg):
    c = np.full(walls_rg.shape, 254, dtype=np.uint8)
    c[walls_rg == 0] = 0
    for k in range(1, 101):
        c[walls_rg == k] = k
    for code, v in ((900, 250), (901, 251), (902, 252), (800, 253), (999, 255)):
        c[walls_rg == code] = v
    return c


def _replay(cfg, walls, seed=3):
    """(device-style result, oracle result, parked) for one stream + bounce-back of random populations."""
    D, Q = cfg.ndims, cfg.Q
    NZ, NY, NX = walls.shape
    R = 1
    Rz = R if D == 3 else 0
    o = oracle.Oracle(cfg)
    o.set_walls(walls)
    ci = o.lattice()["ci"].astype(np.int32)
    opp = o.lattice()["opposites"]
    cls = _classes(geo.ghosted(walls, R, cfg.periodic, D, wall_ghost=True))
    # extended slab: owned planes + Rz ghost planes; positions in ascending (z, y, x) order of the fluid nodes
    ext_fluid = cls[:, R:R + NY, R:R + NX] == 0
    P = np.zeros(ext_fluid.size + 1, dtype=np.uint32)
    P[1:] = np.cumsum(ext_fluid.ravel())
    nstore = int(P[-1])
    fs = ((nstore + 128 + 127) // 128) * 128
    L = _lib()
    cap = 19 * ext_fluid.size
    dst = np.zeros(cap, dtype=np.uint32)
    src = np.zeros(cap, dtype=np.uint32)
    parked = C.c_longlong()
    per = np.array([cfg.periodic[0], cfg.periodic[1], cfg.periodic[2] if D == 3 else 0], dtype=np.int32)
    ip, up = C.POINTER(C.c_int), C.POINTER(C.c_uint32)
    n = L.spec_build(Q, D, np.ascontiguousarray(ci).ctypes.data_as(ip), NX, NY, NZ, R, Rz, per.ctypes.data_as(ip),
                     np.ascontiguousarray(cls).ctypes.data_as(C.POINTER(C.c_uint8)), P.ctypes.data_as(up), fs,
                     dst.ctypes.data_as(up), src.ctypes.data_as(up), cap, C.byref(parked))
    dst, src = dst[:n], src[:n]
    # post-collision populations: random on fluid nodes, 0 on solid nodes (what every state of the step holds)
    rng = np.random.default_rng(seed)
    fstar = rng.uniform(0.01, 1.0, size=(NZ, NY, NX, Q, 1))
    fstar[walls != 0] = 0.0
    # --- the device: push (a solid neighbour bounces back into the node's own opposite slot)
    out = np.full(Q * fs, np.nan)
    N3 = (NX, NY, NZ)
    perx = (cfg.periodic[0], cfg.periodic[1], cfg.periodic[2] if D == 3 else 0)

    def pos(x, y, z):
        return int(P[((z + Rz) * NY + y) * NX + x])

    for z in range(NZ):
        for y in range(NY):
            for x in range(NX):
                if walls[z, y, x] != 0:
                    continue
                here = pos(x, y, z)
                for q in range(Q):
                    cx, cy, cz = (int(v) for v in ci[q])
                    v = fstar[z, y, x, q, 0]
                    if cls[z + Rz + cz, y + R + cy, x + R + cx] != 0:
                        out[opp[q] * fs + here] = v
                    else:
                        t = [x + cx, y + cy, z + cz]
                        for d in range(3):
                            if perx[d]:
                                t[d] %= N3[d]
                        out[q * fs + pos(*t)] = v
    # --- the free-slip slots
    tmp = np.where(src == 0xFFFFFFFF, 0.0, out[np.minimum(src, out.size - 1)])
    out[dst] = tmp
    got = np.zeros((NZ, NY, NX, Q))
    for z in range(NZ):
        for y in range(NY):
            for x in range(NX):
                if walls[z, y, x] == 0:
                    got[z, y, x] = out[np.arange(Q) * fs + pos(x, y, z)]
    # --- the reference's sweep
    o.set_fi(fstar)
    o.phase("communicate_fi")
    o.phase("stream")
    o.phase("bounceback")
    want = o.fi()[..., 0]
    return got, want, parked.value, n


def _cfg(D, NX, NY, NZ, periodic):
    c = tc.default_config(D, 1, NX, NY, NZ)
    for d in range(3):
        c.periodic[d] = periodic[d]
    tc.finalize_flags(c)
    return c


def _check(cfg, walls):
    got, want, parked, n = _replay(cfg, walls)
    assert parked == 0
    assert n > 0
    fluid = walls == 0
    assert np.isfinite(want[fluid]).all()
    assert np.array_equal(got[fluid], want[fluid])


def _disc(walls, cx, cy, r, code=1.0):
    NZ, NY, NX = walls.shape
    yy, xx = np.mgrid[0:NY, 0:NX]
    for z in range(NZ):
        walls[z][(xx - cx) ** 2 + (yy - cy) ** 2 <= r * r] = code


@pytest.mark.parametrize("xper", [1, 0])
def test_duct_2d(xper):
    """initialize_walls_nostick_duct (2-D, y-normal), x periodic or closed by ghost walls, an obstacle inside."""
    NX, NY = 14, 11
    walls = np.zeros((1, NY, NX))
    walls[0, 0, :] = walls[0, -1, :] = tc.WALL_NORMAL_Y
    _disc(walls, 6.5, 5.0, 1.6)
    _check(_cfg(2, NX, NY, 1, (xper, 0, 0)), walls)


def test_duct_2d_x_normal():
    NX, NY = 10, 12
    walls = np.zeros((1, NY, NX))
    walls[0, :, 0] = walls[0, :, -1] = tc.WALL_NORMAL_X
    _disc(walls, 4.5, 6.0, 1.2, code=2.0)
    _check(_cfg(2, NX, NY, 1, (0, 1, 0)), walls)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_duct_3d(axis):
    """initialize_walls_nostick_duct (3-D) for each -duct_normal_direction; the other two axes periodic."""
    N = [9, 8, 7]
    walls = np.zeros((N[2], N[1], N[0]))
    sl = [slice(None)] * 3
    for idx in (0, -1):
        sl[2 - axis] = idx
        walls[tuple(sl)] = 900.0 + axis
    walls[3, 4, 4] = walls[3, 3, 4] = 3.0  # an obstacle away from the mirrors
    per = [1, 1, 1]
    per[axis] = 0
    _check(_cfg(3, N[0], N[1], N[2], per), walls)


def test_wall_row_with_a_gap_and_mixed_codes():
    """A wall row that alternates plain and free-slip nodes and has a fluid gap: several writers compete for
    one slot and the sweep order decides (last writer wins); populations pushed sideways into a mirror are lost."""
    NX, NY = 12, 9
    walls = np.zeros((1, NY, NX))
    walls[0, 0, :] = tc.WALL_NORMAL_Y
    walls[0, 0, 3] = 1.0
    walls[0, 0, 7] = 0.0
    walls[0, -1, :] = 1.0
    cfg = _cfg(2, NX, NY, 1, (1, 0, 0))
    got, want, parked, n = _replay(cfg, walls)
    fluid = walls == 0
    if parked == 0:
        assert np.array_equal(got[fluid], want[fluid])
    else:
        assert parked > 0  # refused by txg_set_walls


def test_corner_of_two_mirrors_is_refused():
    NX, NY = 8, 8
    walls = np.zeros((1, NY, NX))
    walls[0, 0, :] = walls[0, -1, :] = tc.WALL_NORMAL_Y
    walls[0, :, 0] = walls[0, :, -1] = tc.WALL_NORMAL_X
    got, want, parked, n = _replay(_cfg(2, NX, NY, 1, (0, 0, 0)), walls)
    assert parked > 0


def test_obstacle_touching_a_mirror_is_refused():
    NX, NY = 10, 8
    walls = np.zeros((1, NY, NX))
    walls[0, 0, :] = walls[0, -1, :] = tc.WALL_NORMAL_Y
    walls[0, 1, 4] = 1.0
    got, want, parked, n = _replay(_cfg(2, NX, NY, 1, (1, 0, 0)), walls)
    assert parked > 0
