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
