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


def test_load_local_cuts_the_slab_and_fills_periodic_ghosts(tmp_path):
    """IOLoad + DMGlobalToLocal (lbm.F90:482-544): a natural-order fi file becomes each rank's ghosted local array."""
    import cases
    from taxila_lbm_b200 import geometry as geo
    from taxila_lbm_b200 import slab

    cfg, walls, rho = cases.porous_3d(12, NZ=10, rmin=2.0, rmax=3.0, periodic=(1, 0, 1))
    rng = np.random.default_rng(3)
    fi = rng.uniform(size=(10, 12, 12, 19, 2))
    p = tmp_path / petsc_io.output_name("r_", "fi", 4)
    petsc_io.write_vec(p, fi)
    parts = []
    for r in range(3):
        c = slab.local_config(cfg, 3, r)
        loc = petsc_io.load_local(p, c, (19, 2), 1)
        assert loc.shape == (c.zl + 2, 14, 14, 19, 2)
        assert np.array_equal(loc, geo.ghosted(fi, 1, cfg.periodic, 3, zs=c.zs, zl=c.zl))
        assert np.array_equal(loc[:, :, 0], loc[:, :, -2])      # x periodic: low ghost = last owned column
        assert np.all(loc[:, 0] == 0.0)                          # y not periodic: ghost rows untouched
        parts.append(geo.owned(loc, 1, 3))
    assert np.array_equal(np.concatenate(parts, axis=0), fi)
    rp = tmp_path / "rho.dat"
    petsc_io.write_vec(rp, rho)
    assert np.array_equal(geo.owned(petsc_io.load_local(rp, cfg, (2,), cfg.stencil_size_rho), 1, 3), rho)
