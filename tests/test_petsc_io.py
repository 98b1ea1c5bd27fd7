"""The PETSc binary Vec container (the reference's I/O seam, src/lbm/lbm_io.F90:71-93,
src/testing/PetscBinaryRead.py): reader/writer round trip and the shipped golden's header."""
from pathlib import Path

import numpy as np

from taxila_lbm_b200 import petsc_io

GOLDEN = Path(__file__).resolve().parent / "golden" / "bubble_2D_fi001.dat"


def test_golden_is_a_petsc_vec():
    v = petsc_io.read_vec(GOLDEN)
    assert v.size == 128 * 128 * 9 * 2
    assert np.isfinite(v).all()


def test_round_trip(tmp_path):
    rng = np.random.default_rng(7)
    a = rng.standard_normal((5, 4, 3, 2))
    p = tmp_path / petsc_io.output_name("out_", "rho", 7)
    assert p.name == "out_rho007.dat"
    petsc_io.write_vec(p, a)
    b = petsc_io.read_vec(p)
    assert np.array_equal(b, a.ravel())
    # byte-identical to what the reference's reader expects: big-endian header then big-endian doubles
    raw = p.read_bytes()
    assert int.from_bytes(raw[:4], "big") == 1211214 and int.from_bytes(raw[4:8], "big") == a.size


def test_rewriting_the_golden_reproduces_it(tmp_path):
    v = petsc_io.read_vec(GOLDEN)
    p = tmp_path / "fi001.dat"
    petsc_io.write_vec(p, v)
    assert p.read_bytes() == GOLDEN.read_bytes()
