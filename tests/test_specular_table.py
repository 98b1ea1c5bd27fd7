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


# ------------------------------------------------------------------ two z-slabs (the device's sequence, replayed in numpy)
def _replay_two_slabs(cfg, walls, zsplit, seed=5):
    """The multi-rank device sequence on two z-slabs [0, zsplit) and [zsplit, NZ) of a 3-D box, slab by slab:
    push inside the extended slab (pushes across a z face land in ghost-plane positions) -> exchange_f (the crossing
    rows of a ghost plane go to the neighbour's boundary plane, masked: a slot whose streaming source is solid keeps what
    its own node wrote) -> exchange_parked (the boundary planes' rows that point AWAY from the neighbour are copied into the
    neighbour's ghost plane) -> the slab's free-slip table, built with z NOT periodic so that it looks into the ghost
    planes.  Returns (assembled device-style result, oracle result, parked)."""
    D, Q = 3, cfg.Q
    NZ, NY, NX = walls.shape
    R = Rz = 1
    o = oracle.Oracle(cfg)
    o.set_walls(walls)
    ci = o.lattice()["ci"].astype(np.int32)
    opp = o.lattice()["opposites"]
    rng = np.random.default_rng(seed)
    fstar = rng.uniform(0.01, 1.0, size=(NZ, NY, NX, Q, 1))
    fstar[walls != 0] = 0.0
    L = _lib()
    ip, up = C.POINTER(C.c_int), C.POINTER(C.c_uint32)
    slabs = []
    for zs, zl in ((0, zsplit), (zsplit, NZ - zsplit)):
        cls = _classes(geo.ghosted(walls, R, cfg.periodic, D, zs=zs, zl=zl, wall_ghost=True))
        ext_fluid = cls[:, R:R + NY, R:R + NX] == 0
        P = np.zeros(ext_fluid.size + 1, dtype=np.uint32)
        P[1:] = np.cumsum(ext_fluid.ravel())
        fs = ((int(P[-1]) + 128 + 127) // 128) * 128
        cap = 19 * ext_fluid.size
        dst = np.zeros(cap, dtype=np.uint32)
        src = np.zeros(cap, dtype=np.uint32)
        parked = C.c_longlong()
        per = np.array([cfg.periodic[0], cfg.periodic[1], 0], dtype=np.int32)  # several ranks: z looks into the ghost planes
        n = L.spec_build(Q, D, np.ascontiguousarray(ci).ctypes.data_as(ip), NX, NY, zl, R, Rz, per.ctypes.data_as(ip),
                         np.ascontiguousarray(cls).ctypes.data_as(C.POINTER(C.c_uint8)), P.ctypes.data_as(up), fs,
                         dst.ctypes.data_as(up), src.ctypes.data_as(up), cap, C.byref(parked))
        slabs.append(dict(zs=zs, zl=zl, cls=cls, P=P, fs=fs, dst=dst[:n], src=src[:n], parked=parked.value,
                          out=np.full(Q * fs, np.nan)))

    def pos(s, x, y, zloc):  # zloc = -1 .. zl: local plane incl. the ghost planes
        return int(s["P"][((zloc + Rz) * NY + y) * NX + x])

    # push, slab by slab (x, y wrap inside the slab; z never wraps: plane -1 / zl are the ghost planes)
    for s in slabs:
        for zloc in range(s["zl"]):
            z = s["zs"] + zloc
            for y in range(NY):
                for x in range(NX):
                    if walls[z, y, x] != 0:
                        continue
                    here = pos(s, x, y, zloc)
                    for q in range(Q):
                        cx, cy, cz = (int(v) for v in ci[q])
                        v = fstar[z, y, x, q, 0]
                        if s["cls"][zloc + Rz + cz, y + R + cy, x + R + cx] != 0:
                            s["out"][opp[q] * s["fs"] + here] = v
                        else:
                            tx, ty = x + cx, y + cy
                            if cfg.periodic[0]:
                                tx %= NX
                            if cfg.periodic[1]:
                                ty %= NY
                            s["out"][q * s["fs"] + pos(s, tx, ty, zloc + cz)] = v
    # z halos between slab a (below) and slab b (above) across BOTH faces of the periodic box
    a, b = slabs
    faces = [(a, a["zl"] - 1, b, 0)] + ([(b, b["zl"] - 1, a, 0)] if cfg.periodic[2] else [])
    snapshot = [s["out"].copy() for s in slabs]  # every exchange reads what the push left (sends precede receives)
    for lo, zt, hi, zb in faces:  # lo's top plane zt faces hi's bottom plane zb
        lo_snap = snapshot[slabs.index(lo)]
        hi_snap = snapshot[slabs.index(hi)]
        for y in range(NY):
            for x in range(NX):
                zg_lo, zg_hi = lo["zs"] + zt, hi["zs"] + zb
                # exchange_f: lo's top ghost plane (the image of hi's plane zb) -> hi's plane zb, directions c_z > 0
                if walls[zg_hi, y, x] == 0:
                    for q in range(Q):
                        cx, cy, cz = (int(v) for v in ci[q])
                        sx, sy = (x - cx) % NX if cfg.periodic[0] else x - cx, (y - cy) % NY if cfg.periodic[1] else y - cy
                        inside = 0 <= sx < NX and 0 <= sy < NY
                        if cz > 0 and inside and walls[zg_lo, sy, sx] == 0:
                            hi["out"][q * hi["fs"] + pos(hi, x, y, zb)] = lo_snap[q * lo["fs"] + pos(lo, x, y, zt + 1)]
                        # exchange_parked: hi's bottom plane rows c_z > 0 -> lo's top ghost plane
                        if cz > 0:
                            lo["out"][q * lo["fs"] + pos(lo, x, y, zt + 1)] = hi_snap[q * hi["fs"] + pos(hi, x, y, zb)]
                if walls[zg_lo, y, x] == 0:
                    for q in range(Q):
                        cx, cy, cz = (int(v) for v in ci[q])
                        sx, sy = (x - cx) % NX if cfg.periodic[0] else x - cx, (y - cy) % NY if cfg.periodic[1] else y - cy
                        inside = 0 <= sx < NX and 0 <= sy < NY
                        if cz < 0 and inside and walls[zg_hi, sy, sx] == 0:
                            lo["out"][q * lo["fs"] + pos(lo, x, y, zt)] = hi_snap[q * hi["fs"] + pos(hi, x, y, zb - 1)]
                        if cz < 0:
                            hi["out"][q * hi["fs"] + pos(hi, x, y, zb - 1)] = lo_snap[q * lo["fs"] + pos(lo, x, y, zt)]
    got = np.zeros((NZ, NY, NX, Q))
    for s in slabs:
        out = s["out"]
        tmp = np.where(s["src"] == 0xFFFFFFFF, 0.0, out[np.minimum(s["src"], out.size - 1)])
        out[s["dst"]] = tmp
        for zloc in range(s["zl"]):
            z = s["zs"] + zloc
            for y in range(NY):
                for x in range(NX):
                    if walls[z, y, x] == 0:
                        got[z, y, x] = out[np.arange(Q) * s["fs"] + pos(s, x, y, zloc)]
    o.set_fi(fstar)
    o.phase("communicate_fi")
    o.phase("stream")
    o.phase("bounceback")
    return got, o.fi()[..., 0], sum(s["parked"] for s in slabs)


@pytest.mark.parametrize("axis,zper", [(1, 1), (0, 1), (1, 0)])
def test_duct_3d_on_two_slabs(axis, zper):
    """Free-slip walls with normal x or y in a box cut into two z-slabs: reflections with c_z != 0 cross the slab face,
    and an obstacle sits astride it.  The exchanged ghost rows + the per-slab tables reproduce the oracle's sweep bit for bit."""
    N = [9, 8, 10]
    walls = np.zeros((N[2], N[1], N[0]))
    sl = [slice(None)] * 3
    for idx in (0, -1):
        sl[2 - axis] = idx
        walls[tuple(sl)] = 900.0 + axis
    walls[4, 4, 4] = walls[5, 4, 4] = walls[5, 3, 4] = 3.0  # astride the face between planes 4 and 5
    per = [1, 1, zper]
    per[axis] = 0
    cfg = _cfg(3, N[0], N[1], N[2], per)
    got, want, parked = _replay_two_slabs(cfg, walls, zsplit=5)
    assert parked == 0
    fluid = walls == 0
    assert np.array_equal(got[fluid], want[fluid]), float(np.nanmax(np.abs(got[fluid] - want[fluid])))
