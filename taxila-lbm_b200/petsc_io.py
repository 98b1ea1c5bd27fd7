"""PETSc binary Vec container used by every file that crosses the reference's I/O seam
(walls files, fi/rho/u/rhot/prs outputs, goldens): big-endian int32 classid 1211214,
int32 n, n float64 (src/testing/PetscBinaryRead.py:19-24,45-55; written by IOView,
src/lbm/lbm_io.F90:71-93, in DMDA natural ordering)."""
import numpy as np

VEC_CLASSID = 1211214


def read_vec(path):
    with open(path, "rb") as fh:
        hdr = np.frombuffer(fh.read(8), dtype=">i4")
        if hdr[0] != VEC_CLASSID:
            raise ValueError("%s: not a PETSc binary Vec (classid %d)" % (path, hdr[0]))
        data = np.frombuffer(fh.read(8 * int(hdr[1])), dtype=">f8")
        if data.size != hdr[1]:
            raise ValueError("%s: truncated Vec" % path)
    return data.astype(np.float64)


def write_vec(path, array):
    a = np.ascontiguousarray(array, dtype=np.float64).ravel()
    with open(path, "wb") as fh:
        np.array([VEC_CLASSID, a.size], dtype=">i4").tofile(fh)
        a.astype(">f8").tofile(fh)


def output_name(prefix, name, counter):
    """<prefix><name>NNN.dat, three digits (lbm_io.F90:71-83, MAXIODIGITS lbm_definitions.h:21)."""
    return "%s%s%03d.dat" % (prefix, name, counter)


def load_local(path, cfg, dof_shape, width):
    """IOLoad / IOLoadFile + DMGlobalToLocal (lbm.F90:482-544, lbm_io.F90:95-122): read a Vec written in DMDA natural
    ordering ([z][y][x][dofs]) and return this rank's local ghosted array of ghost width `width`
    ([zl+2w][NY+2w][NX+2w] + dof_shape; periodic ghosts filled, others zero) for the slab in cfg (zs, zl)."""
    from . import geometry as geo

    D = cfg.ndims
    NZ = cfg.NZ if D == 3 else 1
    n = (NZ, cfg.NY, cfg.NX) + tuple(dof_shape)
    v = read_vec(path)
    if v.size != int(np.prod(n)):
        raise ValueError("%s holds %d values, the box needs %d" % (path, v.size, int(np.prod(n))))
    zs, zl = (cfg.zs, cfg.zl) if D == 3 else (0, 1)
    return geo.ghosted(v.reshape(n), width, cfg.periodic, D, zs=zs, zl=zl)
