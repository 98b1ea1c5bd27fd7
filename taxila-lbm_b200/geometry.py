"""Synthetic inputs for the flow hot path: the host-side equivalents of the
reference's pluggable `initialize_state` / `initialize_walls` routines and of its
DMDA ghosted local arrays.  Host setup only -- nothing here is on the timed path.

Array conventions (C order == the reference's Fortran arrays with the first index
fastest, lbm_flow.F90:963-970):
  natural   [z][y][x][...dofs]          PETSc global Vec ordering, owned nodes only
  ghosted   [z+w][y+w][x+w][...dofs]    a rank's local array, ghost width w
"""
import numpy as np

from .config import WALL_GHOST


# ------------------------------------------------------------------ initial states
def bubble_rho(cfg, rho_inner, rho_outer, half_width=None):
    """tests/bubble_2D/initialize_state.F90:187-203 (square, half width 26) and
    tests/bubble_3D/initialize_state.F90:104-122 (cube, half width 10): nodes with
    every 1-based index in [(N+1)/2 - hw, (N+1)/2 + hw] get rho_inner, the rest rho_outer."""
    D = cfg.ndims
    if half_width is None:
        half_width = 26 if D == 2 else 10
    NZ = cfg.NZ if D == 3 else 1
    S = cfg.ncomponents
    rho = np.empty((NZ, cfg.NY, cfg.NX, S))
    rho[...] = np.asarray(rho_outer, dtype=np.float64)

    def rng(N):
        c = (N + 1) // 2
        return slice(c - half_width - 1, c + half_width)  # 1-based inclusive -> 0-based slice

    zsl = rng(cfg.NZ) if D == 3 else slice(None)
    rho[zsl, rng(cfg.NY), rng(cfg.NX), :] = np.asarray(rho_inner, dtype=np.float64)
    return rho


def flushing_rho(cfg, walls, rho_invading, rho_defending, axis="z", width=10):
    """src/problem_specs/initialize_state_flushing.F90:97-152: fluid nodes whose 1-based
    index along `axis` is <= 10 get rho_invading, other fluid nodes rho_defending;
    solid nodes keep rho = 0."""
    D = cfg.ndims
    NZ = cfg.NZ if D == 3 else 1
    S = cfg.ncomponents
    rho = np.empty((NZ, cfg.NY, cfg.NX, S))
    rho[...] = np.asarray(rho_defending, dtype=np.float64)
    sl = [slice(None)] * 3
    sl[{"z": 0, "y": 1, "x": 2}[axis]] = slice(0, width)
    rho[tuple(sl)] = np.asarray(rho_invading, dtype=np.float64)
    rho[np.asarray(walls).reshape(NZ, cfg.NY, cfg.NX) != 0] = 0.0
    return rho


# ------------------------------------------------------------------ geometries
def porous_spheres(NX, NY, NZ, seed=20260, rmin=10.0, rmax=22.0, solid_fraction=0.55, nminerals=3, periodic=True):
    """Random overlapping spheres: numpy default_rng(seed);
    per sphere draw radius ~ U[rmin, rmax] then centre ~ U[0,N) per axis (x, y, z order);
    voxels with squared distance (periodic minimum image) <= r^2 become solid with mineral id
    1 + (k mod nminerals), k the sphere index (a later sphere overwrites an earlier id);
    spheres are added until the solid fraction is >= solid_fraction.
    Returns walls as float64 [z][y][x] holding the mineral id (0 = pore)."""
    rng = np.random.default_rng(seed)
    walls = np.zeros((NZ, NY, NX), dtype=np.float64)
    total = NX * NY * NZ
    nsolid = 0
    k = 0
    while nsolid < solid_fraction * total:
        r = rng.uniform(rmin, rmax)
        cx, cy, cz = rng.uniform(0, NX), rng.uniform(0, NY), rng.uniform(0, NZ)
        ir = int(np.ceil(r))
        xs = np.arange(int(np.floor(cx)) - ir, int(np.floor(cx)) + ir + 2)
        ys = np.arange(int(np.floor(cy)) - ir, int(np.floor(cy)) + ir + 2)
        zs = np.arange(int(np.floor(cz)) - ir, int(np.floor(cz)) + ir + 2)
        if NZ == 1:
            zs = np.array([0])
            cz = 0.0
        d2 = ((zs - cz) ** 2)[:, None, None] + ((ys - cy) ** 2)[None, :, None] + ((xs - cx) ** 2)[None, None, :]
        inside = d2 <= r * r
        if periodic:
            zi, yi, xi = zs % NZ, ys % NY, xs % NX
        else:
            keep_z, keep_y, keep_x = (zs >= 0) & (zs < NZ), (ys >= 0) & (ys < NY), (xs >= 0) & (xs < NX)
            inside = inside[keep_z][:, keep_y][:, :, keep_x]
            zi, yi, xi = zs[keep_z], ys[keep_y], xs[keep_x]
        # a box wider than the domain would alias under modulo indexing; radii << N here
        assert len(np.unique(zi)) == len(zi) and len(np.unique(yi)) == len(yi) and len(np.unique(xi)) == len(xi)
        sub = walls[np.ix_(zi, yi, xi)]
        nsolid += int(np.count_nonzero(inside & (sub == 0)))
        sub[inside] = 1 + (k % nminerals)
        walls[np.ix_(zi, yi, xi)] = sub
        k += 1
    return walls


def duct_walls(NX, NY, NZ=1, code=1.0):
    """src/problem_specs/initialize_walls_duct.F90:57-72 flavour: solid layers on the low and
    high faces of every non-flow direction (y in 2-D; y and z in 3-D), flow along x."""
    walls = np.zeros((NZ, NY, NX))
    walls[:, 0, :] = code
    walls[:, -1, :] = code
    if NZ > 1:
        walls[0, :, :] = code
        walls[-1, :, :] = code
    return walls


def random_walls(NX, NY, NZ=1, solid_fraction=0.3, seed=1, nminerals=1):
    """Uncorrelated random solid voxels (the survey's 30 %-solid bounce-back check)."""
    rng = np.random.default_rng(seed)
    solid = rng.random((NZ, NY, NX)) < solid_fraction
    ids = rng.integers(1, nminerals + 1, size=(NZ, NY, NX))
    return np.where(solid, ids, 0).astype(np.float64)


# ------------------------------------------------------------------ ghosted local arrays
def ghosted(nat, width, periodic, ndims, zs=0, zl=None, fill=0.0, wall_ghost=False):
    """A rank's local ghosted array from the global natural-order array `nat`
    ([z][y][x][...]).  Periodic directions wrap (DM_BOUNDARY_PERIODIC); ghosts beyond a
    non-periodic face hold `fill`, or WALL_GHOST when wall_ghost (WallsSetGhostNodes,
    lbm_walls.F90:190-231).  The slab [zs, zs+zl) is cut along z (3-D only)."""
    nat = np.asarray(nat)
    NZ, NY, NX = nat.shape[:3]
    if zl is None:
        zl = NZ
    w = width
    wz = w if ndims == 3 else 0
    fillv = WALL_GHOST if wall_ghost else fill

    def idx(start, length, N, per, wd):
        g = np.arange(start - wd, start + length + wd)
        out = (g < 0) | (g >= N)
        return (g % N), (out if not per else np.zeros_like(out))

    zi, zo = idx(zs, zl, NZ, bool(periodic[2]) if ndims == 3 else True, wz)
    yi, yo = idx(0, NY, NY, bool(periodic[1]), w)
    xi, xo = idx(0, NX, NX, bool(periodic[0]), w)
    loc = nat[np.ix_(zi, yi, xi)].copy()
    loc[zo, :, :] = fillv
    loc[:, yo, :] = fillv
    loc[:, :, xo] = fillv
    return np.ascontiguousarray(loc)


def owned(loc, width, ndims):
    """Strip the ghost layers of a local ghosted array."""
    w = width
    wz = w if ndims == 3 else 0
    nz, ny, nx = loc.shape[:3]
    return loc[wz : nz - wz, w : ny - w, w : nx - w]
