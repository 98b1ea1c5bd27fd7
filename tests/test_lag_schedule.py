"""The block schedule of the one-pass step (csrc/lag_schedule.h, compiled here with g++) without a GPU: every
owned position is collided by exactly one C block, every position of planes 1 .. NZl-2 is summed by exactly one
M block, and -- against the real adjacency of random porous boxes -- every node that pushes into a position of
an M block belongs to a C block of one of that M block's dependency rows, all of whose C blocks come earlier in
the linear launch order (the no-deadlock / completeness claim in the header)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent
SRC = HERE / "native" / "lag_schedule_capi.cpp"
HDR = HERE.parent / "taxila-lbm_b200" / "csrc" / "lag_schedule.h"
LIB = HERE / "native" / "_build" / "liblag_schedule.so"
MAX_BANDS = 16
D3Q19 = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (-1, 0, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1), (1, 1, 0), (-1, 1, 0), (-1, -1, 0),
         (1, -1, 0), (1, 0, 1), (-1, 0, 1), (-1, 0, -1), (1, 0, -1), (0, 1, 1), (0, -1, 1), (0, -1, -1), (0, 1, -1)]


def _lib():
    LIB.parent.mkdir(exist_ok=True)
    if not LIB.exists() or LIB.stat().st_mtime < max(SRC.stat().st_mtime, HDR.stat().st_mtime):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-shared", "-fPIC", "-o", str(LIB), str(SRC)],
                       check=True)
    L = C.CDLL(str(LIB))
    up, ip, lp = C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_longlong)
    L.lag_build.restype = C.c_longlong
    L.lag_build.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, up, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, up, C.c_longlong,
                            lp, ip]
    return L


def build(fluid_ext, pery, PB, MB, BR, lag, max_rows=7680):
    """fluid_ext: [NZl + 2][NY][NX] bool (one ghost plane each side).  Returns the schedule as the kernel sees it:
    rows [nrows][6], (nbands, rows_per_band, grid_x, c_blocks, m_blocks), depbands [16][3], and P."""
    nzE, NY, NX = fluid_ext.shape
    P = np.zeros(fluid_ext.size + 1, dtype=np.uint32)
    P[1:] = np.cumsum(fluid_ext.ravel())
    row_off = np.ascontiguousarray(P[::NX][: nzE * NY + 1])
    assert row_off.size == nzE * NY + 1
    cap = 4 * nzE * NY
    rows = np.zeros(cap * 6, dtype=np.uint32)
    meta = np.zeros(5, dtype=np.int64)
    depbands = np.zeros(MAX_BANDS * 3, dtype=np.int32)
    up, ip, lp = C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_longlong)
    n = _lib().lag_build(NY, nzE - 2, 1, int(pery), row_off.ctypes.data_as(up), PB, MB, BR, lag, max_rows,
                         rows.ctypes.data_as(up), cap, meta.ctypes.data_as(lp), depbands.ctypes.data_as(ip))
    if n < 0:
        return None
    assert n <= cap
    return rows[: n * 6].reshape(n, 6).astype(np.int64), meta, depbands.reshape(MAX_BANDS, 3), P


def tasks_of(rows, meta, depbands, PB, MB, lag):
    """The blocks in linear launch order, decoded exactly as k_step_fused_lag decodes its block index (row = index / grid_x, x = index % grid_x):
    (linear index, first, count, is_m, row, dependency rows)."""
    rpb, gx = int(meta[1]), int(meta[2])
    nblk = lambda c, per: (c + per - 1) // per  # noqa: E731
    out = []
    for y in range(rows.shape[0]):
        cfirst, ccount, m0f, m0c, m1f, m1c = (int(v) for v in rows[y])
        nc, n0, n1 = nblk(ccount, PB), nblk(m0c, MB), nblk(m1c, MB)
        assert nc + n0 + n1 <= gx
        b, k = divmod(y, rpb)
        zm = k - 1 - lag
        deps = [int(bb) * rpb + zm + dz for bb in depbands[b] if bb >= 0 for dz in (-1, 0, 1)]
        for x in range(nc + n0 + n1):
            lin = y * gx + x
            if x < nc:
                out.append((lin, cfirst + x * PB, min(PB, ccount - x * PB), 0, y, ()))
            elif x - nc < n0:
                q = x - nc
                out.append((lin, m0f + q * MB, min(MB, m0c - q * MB), 1, y, deps))
            else:
                q = x - nc - n0
                out.append((lin, m1f + q * MB, min(MB, m1c - q * MB), 1, y, deps))
    return out


def check(NX, NY, NZl, solid_fraction, pery, PB, MB, BR, lag, seed, max_rows=7680):
    rng = np.random.default_rng(seed)
    owned = rng.random((NZl, NY, NX)) >= solid_fraction
    owned[:, :, 0] |= rng.random((NZl, NY)) < 0.5  # some rows start fluid, some do not
    fluid_ext = np.concatenate([owned[-1:], owned, owned[:1]])  # periodic z on one rank: ghosts mirror the far planes
    rows, meta, depbands, P = build(fluid_ext, pery, PB, MB, BR, lag, max_rows)
    assert rows.shape[0] == meta[0] * meta[1] and rows.shape[0] <= max_rows and meta[0] <= MAX_BANDS
    tasks = tasks_of(rows, meta, depbands, PB, MB, lag)
    assert sum(1 for t in tasks if not t[3]) == meta[3] and sum(1 for t in tasks if t[3]) == meta[4]
    nstore = int(P[-1])
    plane = NY * NX
    own0, own1 = int(P[plane]), int(P[(NZl + 1) * plane])
    m0, m1 = int(P[2 * plane]), int(P[NZl * plane])  # positions of owned planes 1 .. NZl-2
    # coverage
    ccover = np.zeros(nstore, dtype=np.int64)
    mcover = np.zeros(nstore, dtype=np.int64)
    crow = np.full(nstore, -1, dtype=np.int64)  # schedule row of the C block that collides a position
    clin = np.full(nstore, -1, dtype=np.int64)  # its linear block index
    for lin, first, count, is_m, y, deps in tasks:
        sl = slice(first, first + count)
        assert 1 <= count <= (MB if is_m else PB)
        if is_m:
            mcover[sl] += 1
        else:
            ccover[sl] += 1
            crow[sl] = y
            clin[sl] = lin
    assert np.all(ccover[own0:own1] == 1) and ccover[:own0].sum() == 0 and ccover[own1:].sum() == 0
    assert np.all(mcover[m0:m1] == 1) and mcover[:m0].sum() == 0 and mcover[m1:].sum() == 0
    last_c = np.full(rows.shape[0], -1, dtype=np.int64)  # the last linear index of the C blocks of every row
    for lin, first, count, is_m, y, deps in tasks:
        if not is_m:
            last_c[y] = max(last_c[y], lin)
    # dependency completeness against the real adjacency
    pos_to_node = np.flatnonzero(fluid_ext.ravel())
    for lin, first, count, is_m, y, deps in tasks:
        if not is_m:
            continue
        d = set(deps)
        assert len(deps) <= 9 and all(0 <= r < rows.shape[0] for r in d)
        for r in d:
            assert last_c[r] < lin, "a dependency is launched after its M block"
        pos = np.arange(first, first + count)
        node = pos_to_node[pos]
        zz, rr = np.divmod(node, plane)
        yy, x = np.divmod(rr, NX)
        assert zz.min() >= 2 and zz.max() <= NZl - 1
        for (cx, cy, cz) in D3Q19:
            sx, sy, sz = (x - cx) % NX, yy - cy, zz - cz
            inside = (sy >= 0) & (sy < NY)
            if pery:
                sy = sy % NY
                inside[:] = True
            sy_c = np.clip(sy, 0, NY - 1)
            src_fluid = inside & fluid_ext[sz, sy_c, sx]
            pusher = np.where(src_fluid, P[(sz * NY + sy_c) * NX + sx], pos)  # solid source: the node's own bounce-back
            assert np.all(crow[pusher] >= 0)
            bad = [int(g) for g in np.unique(crow[pusher]) if int(g) not in d]
            assert not bad, (lin, (cx, cy, cz), bad, d)
            assert np.all(clin[pusher] < lin)
    return len(tasks)


@pytest.mark.parametrize("pery", [1, 0])
@pytest.mark.parametrize("BR,lag", [(2, 0), (5, 1), (8, 2), (64, 1)])
def test_schedule_covers_and_orders(pery, BR, lag):
    check(NX=19, NY=23, NZl=9, solid_fraction=0.5, pery=pery, PB=16, MB=64, BR=BR, lag=lag, seed=BR + lag)


def test_schedule_no_solids_and_single_band():
    check(NX=8, NY=8, NZl=6, solid_fraction=0.0, pery=1, PB=64, MB=512, BR=8, lag=1, seed=1)
    check(NX=8, NY=12, NZl=6, solid_fraction=0.0, pery=1, PB=40, MB=120, BR=4, lag=0, seed=2)


def test_schedule_mostly_solid():
    check(NX=16, NY=16, NZl=8, solid_fraction=0.93, pery=1, PB=16, MB=32, BR=4, lag=1, seed=4)


def test_band_height_grows_to_fit_the_row_budget():
    """max_rows (the constant-memory table of the kernel) and LAG_MAX_BANDS bound the number of bands."""
    check(NX=8, NY=40, NZl=6, solid_fraction=0.3, pery=1, PB=16, MB=64, BR=2, lag=1, seed=5)  # 20 bands -> <= 16
    check(NX=8, NY=40, NZl=6, solid_fraction=0.3, pery=1, PB=16, MB=64, BR=4, lag=1, seed=6, max_rows=30)  # <= 4 bands


def test_thin_boxes_are_not_eligible():
    fl = np.ones((5, 8, 8), dtype=bool)
    assert build(fl, 1, 64, 512, 4, 1) is None  # NZl = 3
