"""z-slab decomposition of a D3Q19 box across ranks (one rank == one GPU).

Replaces the DMDA decomposition of the reference (src/lbm/lbm_grid.F90:159-212) for the
hot path: x and y are never split (run the reference with -da_processors_x 1
-da_processors_y 1), rank r owns a contiguous run of z-planes.  PETSc's DMDA splits N
planes over P ranks as N // P each with the first N % P ranks taking one more
(DMSetUp_DA_3D's default ownership ranges); `slab_range` reproduces that, so a Fortran
driver's local arrays line up with the slabs used here.
"""
import numpy as np

from . import geometry as geo


def slab_range(NZ, nranks, rank):
    """(zs, zl): 0-based first plane and number of planes owned by `rank`."""
    if not (0 <= rank < nranks):
        raise ValueError("rank %d of %d" % (rank, nranks))
    if nranks > NZ:
        raise ValueError("more ranks (%d) than z-planes (%d)" % (nranks, NZ))
    base, extra = divmod(NZ, nranks)
    zl = base + (1 if rank < extra else 0)
    zs = rank * base + min(rank, extra)
    return zs, zl


def neighbours(nranks, rank, periodic_z):
    """(down, up) ranks of the slab ring, -1 where a non-periodic box ends."""
    up = rank + 1 if rank + 1 < nranks else (0 if periodic_z else -1)
    down = rank - 1 if rank > 0 else (nranks - 1 if periodic_z else -1)
    return down, up


def local_config(cfg, nranks, rank):
    """Copy of the global config with this rank's slab filled in."""
    c = cfg.copy()
    c.zs, c.zl = slab_range(cfg.NZ, nranks, rank)
    c.rank, c.nranks = rank, nranks
    return c


def local_arrays(cfg, walls, rho, nranks, rank):
    """This rank's (cfg, walls_rg, rho_rg): the local ghosted arrays a DMDA rank holds after
    WallsSetValues/WallsCommunicate and LBMInitializeState, cut from the global natural arrays."""
    c = local_config(cfg, nranks, rank)
    R = cfg.stencil_size_rho
    walls_rg = geo.ghosted(walls, R, cfg.periodic, 3, zs=c.zs, zl=c.zl, wall_ghost=True)
    rho_rg = geo.ghosted(rho, R, cfg.periodic, 3, zs=c.zs, zl=c.zl)
    return c, walls_rg, rho_rg


def local_bc_values(cfg, bcs, nranks, rank):
    """This rank's face arrays (BCSetUp, lbm_bc.F90:127-213: xm..yp hold the rank's own zs:ze range, zm lives on
    the rank with zs == 1 and zp on the one with ze == NZ) cut from the global ones {boundary: array}."""
    zs, zl = slab_range(cfg.NZ, nranks, rank)
    out = {}
    for b, v in bcs.items():
        if b < 4:
            out[b] = np.ascontiguousarray(v[zs:zs + zl])
        elif (b == 4 and zs == 0) or (b == 5 and zs + zl == cfg.NZ):
            out[b] = v
    return out


def assemble(parts):
    """Global natural-order array from the per-rank owned arrays (rank order == z order)."""
    return np.concatenate(parts, axis=0)
