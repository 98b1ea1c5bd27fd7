"""Pin the CPU oracle to the reference's own golden vector.

tests/golden/bubble_2D_fi001.dat is a byte copy of the reference's
tests/bubble_2D/reference_solution/fi001.dat: full distribution function after 100
steps of the shipped D2Q9 two-component bubble (the file `make test` compares with
src/testing/check_solution.py at eps = 1e-5)."""
import numpy as np
import pytest

import cases


def test_bubble_2d_golden_srt():
    cfg, walls, rho = cases.bubble_2d()
    o = cases.run_oracle(cfg, walls, rho, 100)
    g = cases.golden_bubble_2d()
    err = np.abs(o.fi() - g).max()
    assert err <= 1e-14, err
    # per-component mass of the shipped case
    np.testing.assert_allclose(o.fi().sum(axis=(0, 1, 2, 3)), [13252.02, 3131.98], rtol=1e-12)


def test_bubble_2d_golden_mrt_unit_rates():
    """MRT with every rate 1 is SRT with tau 1: pins the D2Q9 moment matrix and norms."""
    cfg, walls, rho = cases.bubble_2d(mrt=True)
    o = cases.run_oracle(cfg, walls, rho, 100)
    err = np.abs(o.fi() - cases.golden_bubble_2d()).max()
    assert err <= 1e-13, err


def _project_d3q19_to_d2q9(fi3, lat3, lat2):
    """Sum the D3Q19 populations sharing (c_x, c_y) onto the D2Q9 direction with that vector."""
    out = np.zeros(fi3.shape[:3] + (9, fi3.shape[-1]))
    for n in range(19):
        cx, cy = lat3["ci"][n][0], lat3["ci"][n][1]
        k = next(q for q in range(9) if lat2["ci"][q][0] == cx and lat2["ci"][q][1] == cy)
        out[..., k, :] += fi3[..., n, :]
    return out


def test_d3q19_extrusion_projects_onto_2d_golden():
    """A z-invariant D3Q19 run of the same bubble projects exactly (in exact arithmetic) onto
    the D2Q9 golden: pins the D3Q19 tables, Equilf_D3Q19 and the D3 iso-4 force."""
    import oracle

    cfg, walls, rho3 = cases.bubble_3d(N=128, NZ=4, hw=26)
    rho3[:] = cases.bubble_2d()[2][0][None]  # same square in every z plane
    o = cases.run_oracle(cfg, walls, rho3, 100)
    c2 = cases.bubble_2d()[0]
    lat2 = oracle.Oracle(c2).lattice()
    proj = _project_d3q19_to_d2q9(o.fi(), o.lattice(), lat2)
    g = cases.golden_bubble_2d()[0]
    for z in range(4):
        err = np.abs(proj[z] - g).max()
        assert err <= 1e-13, (z, err)
    assert np.abs(o.u()[..., 2, :]).max() <= 1e-15  # no z velocity


def test_d3q19_mrt_unit_rates_equals_srt():
    cfg, walls, rho = cases.bubble_3d(N=24, hw=5)
    a = cases.run_oracle(cfg, walls, rho, 20).fi()
    cfg2, _, _ = cases.bubble_3d(N=24, hw=5, mrt=True)
    b = cases.run_oracle(cfg2, walls, rho, 20).fi()
    assert np.abs(a - b).max() <= 1e-13


def test_mrt_rows_orthogonal():
    import oracle

    for cfg in (cases.bubble_2d(16)[0], cases.bubble_3d(8)[0]):
        lat = oracle.Oracle(cfg).lattice()
        gram = lat["mt"] @ lat["mt"].T
        assert np.array_equal(gram, np.diag(lat["mmt"]))
        assert abs(lat["weights"].sum() - 1.0) < 1e-15


@pytest.mark.parametrize("stem", ["c3_bubble3d_256_1000", "c4_porous_256_1000"])
def test_large_golden_samples_are_well_formed(stem):
    """The committed samples of the oracle's 256^3 x 1000-step runs (tests/golden/make_large_golden.py) that the GPU parity tests read:
    shapes, finiteness, sorted unique node indices inside the box, positive mass."""
    import json
    from pathlib import Path

    path = Path(__file__).resolve().parent / "golden" / (stem + ".npz")
    if not path.exists():
        pytest.skip("sample not generated")
    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    nodes = int(np.prod(meta["box"]))
    idx = g["idx"]
    assert meta["steps"] == 1000 and idx.ndim == 1 and np.all(np.diff(idx) > 0) and 0 <= idx[0] and idx[-1] < nodes
    assert g["rho"].shape == (idx.size, 2) and g["u"].shape[0] == idx.size and g["fi"].shape[1] == 38
    for k in ("rho", "u", "fi", "mass"):
        assert np.isfinite(g[k]).all(), k
    assert (g["mass"] > 0).all() and (g["rho"] >= 0).all()
