"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's golden.

Tolerances: populations vs the shipped golden <= 1e-12 max abs;
rho, u, forces vs the oracle <= 1e-10 * max|field|; node classes byte-identical; per-component
mass conserved to <= 1e-12 relative.
"""
import numpy as np
import pytest

import cases
import gpu_util
from taxila_lbm_b200 import config as tc
from taxila_lbm_b200 import geometry as geo

pytestmark = pytest.mark.gpu

TOL = 1e-10


def compare(cfg, walls, rho, steps, tol=TOL, check_fi=True):
    o = cases.run_oracle(cfg, walls, rho, steps)
    flow = gpu_util.make_flow(cfg, walls, rho)
    flow.step(steps)
    fi, r, u, F = gpu_util.fields(flow)
    fluid = np.asarray(walls).reshape(r.shape[:3]) == 0
    errs = {
        "rho": gpu_util.rel_err(r[fluid], o.rho()[fluid]),
        "u": gpu_util.rel_err(u[fluid], o.u()[fluid]),
        "forces": gpu_util.rel_err(F[fluid], o.forces()[fluid]) if np.abs(o.forces()).max() > 0 else 0.0,
    }
    if check_fi:
        errs["fi"] = gpu_util.rel_err(fi, o.fi())
    for k, v in errs.items():
        assert v <= tol, (k, v, errs)
    # solid nodes hold nothing
    assert np.all(fi[~fluid] == 0.0)
    assert np.all(r[~fluid] == 0.0)
    # mass conservation against the initial state
    m0 = gpu_util.mass(np.asarray(rho).reshape(r.shape), fluid)
    m1 = gpu_util.mass(r, fluid)
    assert np.all(np.abs(m1 - m0) <= 1e-12 * np.abs(m0)), (m0, m1)
    flow.close()
    return errs


def test_bubble_2d_golden():
    cfg, walls, rho = cases.bubble_2d()
    flow = gpu_util.make_flow(cfg, walls, rho)
    flow.step(100)
    fi = geo.owned(flow.get_fi(), 1, 2)
    err = np.abs(fi - cases.golden_bubble_2d()).max()
    assert err <= 1e-12, err


def test_bubble_2d_vs_oracle():
    compare(*cases.bubble_2d(), steps=100)


def test_bubble_2d_mrt_iso8():
    compare(*cases.bubble_2d(64, mrt=True, order=8, hw=12), steps=60)


def test_bubble_2d_hots_mrt_iso10():
    compare(*cases.bubble_2d_hots(), steps=100)


def test_bubble_3d_srt():
    compare(*cases.bubble_3d(32, hw=6), steps=50)


def test_bubble_3d_mrt_iso8():
    compare(*cases.bubble_3d(24, mrt=True, order=8, hw=5), steps=30)


def test_d3q19_extrusion_projects_onto_2d_golden():
    cfg, walls, rho3 = cases.bubble_3d(N=128, NZ=4, hw=26)
    rho3[:] = cases.bubble_2d()[2][0][None]
    flow = gpu_util.make_flow(cfg, walls, rho3)
    flow.step(100)
    fi = geo.owned(flow.get_fi(), 1, 3)
    ci3 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [1, 1, 0], [-1, 1, 0],
                    [-1, -1, 0], [1, -1, 0], [1, 0, 1], [-1, 0, 1], [-1, 0, -1], [1, 0, -1], [0, 1, 1], [0, -1, 1],
                    [0, -1, -1], [0, 1, -1]])
    ci2 = [(0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1)]
    proj = np.zeros(fi.shape[:3] + (9, 2))
    for n in range(19):
        proj[..., ci2.index((ci3[n][0], ci3[n][1])), :] += fi[..., n, :]
    g = cases.golden_bubble_2d()[0]
    assert np.abs(proj - g[None]).max() <= 1e-12


@pytest.mark.parametrize("order", [4, 8])
def test_porous_mrt_minerals_body(order):
    compare(*cases.porous_3d(40, order=order, rmin=4.0, rmax=9.0), steps=100)


def test_porous_split_path_order4(monkeypatch):
    """The order-4 stencil through the split k_forces + k_collide pair (the default is the fused kernel)."""
    monkeypatch.setenv("TXG_SPLIT", "1")
    compare(*cases.porous_3d(40, order=4, rmin=4.0, rmax=9.0), steps=100)


def test_bubble_2d_split_path(monkeypatch):
    monkeypatch.setenv("TXG_SPLIT", "1")
    compare(*cases.bubble_2d(), steps=100)


def test_porous_1000_steps():
    """The parity bar of SURVEY.md 8d: fields within 1e-10 of the oracle after 1000 steps (C4 recipe, 48^3 crop)."""
    compare(*cases.porous_3d(48, order=4, rmin=4.0, rmax=9.0), steps=1000)


@pytest.mark.parametrize("split", [False, True])
def test_porous_shan_chen_eos(monkeypatch, split):
    """-flow_use_nonideal_eos with EOS_SC psi = rho0 (1 - exp(-rho / rho0)) (lbm_eos.F90:183-229): psi
    replaces rho in the fluid-fluid stencil only; fused and split kernels."""
    if split:
        monkeypatch.setenv("TXG_SPLIT", "1")
    cfg, walls, rho = cases.porous_3d(32, order=4, rmin=4.0, rmax=8.0)
    cfg.use_nonideal_eos = 1
    for m in range(2):
        cfg.eos_type[m] = tc.EOS_SC
        cfg.eos_rho0[m] = 0.8 + 0.3 * m
    compare(cfg, walls, rho, steps=60)


def test_three_components_mrt():
    """S = 3: 10 nodes x 3 components per warp (two spare lanes), 3 x 3 coupling matrix, minerals."""
    cfg = tc.default_config(3, 3, 24, 24, 24)
    for d in range(3):
        cfg.periodic[d] = 1
    cfg.relaxation_mode = tc.RELAXATION_MODE_MRT
    for m in range(3):
        cfg.s_e[m], cfg.s_e2[m], cfg.s_q[m], cfg.s_pi[m], cfg.s_m[m] = 1.19, 1.4, 1.2, 1.4, 1.98
        cfg.mm[m] = 1.0 + 0.25 * m
        for k in range(3):
            if k != m:
                cfg.gf[m][k] = 0.05 + 0.01 * (m + k)
    cfg.nminerals = 2
    for k in range(2):
        for m in range(3):
            cfg.gw[k][m] = 0.01 * (k + 1) * (m - 1)
    cfg.body_forces = 1
    cfg.gvt[2] = 1e-5
    tc.finalize_flags(cfg)
    walls = geo.porous_spheres(24, 24, 24, seed=11, rmin=3.0, rmax=6.0, solid_fraction=0.4, nminerals=2)
    rng = np.random.default_rng(5)
    rho = 0.2 + 0.6 * rng.random((24, 24, 24, 3))
    rho[walls != 0] = 0.0
    compare(cfg, walls, rho, steps=50)


@pytest.mark.parametrize("S,ndims,order,mrt", [(4, 3, 4, True), (5, 3, 4, True), (5, 3, 8, False), (4, 2, 10, True), (5, 2, 4, False)])
def test_four_and_five_components(S, ndims, order, mrt):
    """NMAX_COMPONENTS = 5 (lbm_definitions.h:71): 8 nodes x 4 components and 6 nodes x 5 components per warp (two spare
    lanes), full coupling matrix, unequal molecular masses, minerals; fused (order 4) and split (orders 8, 10) paths."""
    N = 20 if ndims == 3 else 40
    cfg = tc.default_config(ndims, S, N, N, N if ndims == 3 else 1)
    for d in range(ndims):
        cfg.periodic[d] = 1
    cfg.isotropy_order = order
    cfg.stencil_size_rho = {4: 1, 8: 2, 10: 3}[order]
    cfg.relaxation_mode = tc.RELAXATION_MODE_MRT if mrt else tc.RELAXATION_MODE_SRT
    for m in range(S):
        cfg.tau[m] = 0.9 + 0.05 * m
        cfg.s_e[m], cfg.s_e2[m], cfg.s_q[m], cfg.s_pi[m], cfg.s_m[m] = 1.19, 1.4, 1.2, 1.4, 1.98
        cfg.mm[m] = 1.0 + 0.2 * m
        for k in range(S):
            if k != m:
                cfg.gf[m][k] = 0.03 + 0.005 * (m + k)
    cfg.nminerals = 2
    for k in range(2):
        for m in range(S):
            cfg.gw[k][m] = 0.01 * (k + 1) * (m - 1.5)
    cfg.body_forces = 1
    cfg.gvt[ndims - 1] = 1e-5
    tc.finalize_flags(cfg)
    NZ = N if ndims == 3 else 1
    walls = geo.porous_spheres(N, N, NZ, seed=13, rmin=2.5, rmax=5.0, solid_fraction=0.35, nminerals=2)
    rho = 0.15 + 0.5 * np.random.default_rng(S).random((NZ, N, N, S))
    rho[walls != 0] = 0.0
    compare(cfg, walls, rho, steps=30)


def test_porous_srt_nonperiodic_box():
    """Closed box: every face non-periodic -> 999 ghosts, bounce-back off the ghost layer."""
    compare(*cases.porous_3d(32, mrt=False, rmin=4.0, rmax=8.0, periodic=(0, 0, 0)), steps=60)


def test_porous_2d_walls_fluidsolid():
    cfg, _, _ = cases.bubble_2d(64, hw=10)
    cfg.nminerals = 2
    cfg.gw[0][0], cfg.gw[0][1] = -0.03, 0.03
    cfg.gw[1][0], cfg.gw[1][1] = 0.02, -0.02
    cfg.body_forces = 1
    cfg.gvt[0] = 1e-5
    tc.finalize_flags(cfg)
    walls = geo.porous_spheres(64, 64, 1, seed=3, rmin=3.0, rmax=7.0, solid_fraction=0.3, nminerals=2)
    rho = geo.bubble_rho(cfg, (0.03, 0.97), (0.97, 0.03), 10)
    rho[walls != 0] = 0.0
    compare(cfg, walls, rho, steps=80)


def test_single_component_srt():
    cfg = tc.default_config(3, 1, 24, 20, 16)
    for d in range(3):
        cfg.periodic[d] = 1
    cfg.body_forces = 1
    cfg.gvt[0] = 1e-4
    cfg.tau[0] = 0.8
    tc.finalize_flags(cfg)
    walls = geo.duct_walls(24, 20, 16)
    rho = np.ones((16, 20, 24, 1))
    rho[walls != 0] = 0.0
    compare(cfg, walls, rho, steps=60)


def test_node_class_bit_exact():
    """walls(rg..) -> device classes -> back must reproduce the oracle's ghosted walls array."""
    import oracle

    cfg, walls, rho = cases.porous_3d(24, rmin=3.0, rmax=6.0, periodic=(1, 0, 1), order=8)
    walls[3, 4, 5] = 800.0
    walls[0, 0, 0] = 77.0
    o = oracle.Oracle(cfg)
    o.set_walls(walls)
    ref = o.walls_rg()
    flow = gpu_util.make_flow(cfg, walls, rho)
    cls = flow.node_class()
    back = cls.astype(np.float64)
    back[cls == 253] = 800.0
    back[cls == 255] = 999.0
    assert np.array_equal(back, ref)
    assert np.array_equal(cls == 0, ref == 0.0)


def test_set_fi_get_fi_roundtrip_exact():
    cfg, walls, rho = cases.porous_3d(24, rmin=3.0, rmax=6.0)
    flow = gpu_util.make_flow(cfg, walls, rho)
    flow.step(7)
    fi = flow.get_fi()
    flow2 = gpu_util.make_flow(cfg, walls, rho)
    fi_in = geo.ghosted(geo.owned(fi, 1, 3), 1, cfg.periodic, 3)
    flow2.set_fi(fi_in)
    assert np.array_equal(geo.owned(flow2.get_fi(), 1, 3), geo.owned(fi, 1, 3))
    flow.step(5)
    flow2.step(5)
    assert np.array_equal(flow2.get_fi(), flow.get_fi())


def test_six_procedure_sequence_equals_step():
    cfg, walls, rho = cases.bubble_3d(16, hw=3)
    a = gpu_util.make_flow(cfg, walls, rho)
    b = gpu_util.make_flow(cfg, walls, rho)
    a.step(3)
    for _ in range(3):
        b.collision(); b.communicate_fi(); b.stream(); b.bounceback(); b.apply_bcs(); b.update_flux()
    assert np.array_equal(a.get_fi(), b.get_fi())
    from taxila_lbm_b200.capi import TaxilaGpuError

    with pytest.raises(TaxilaGpuError):
        b.stream()  # out of order


def test_diagnostics_vs_oracle():
    cfg, walls, rho = cases.porous_3d(32, rmin=4.0, rmax=8.0)
    o = cases.run_oracle(cfg, walls, rho, 40)
    flow = gpu_util.make_flow(cfg, walls, rho)
    flow.step(40)
    rhot, prs, velt = flow.update_diagnostics()
    orhot, oprs, ovelt = o.diagnostics()
    assert gpu_util.rel_err(rhot, orhot) <= TOL
    assert gpu_util.rel_err(prs, oprs) <= TOL
    assert gpu_util.rel_err(velt, ovelt) <= TOL


@pytest.mark.parametrize("case", ["porous", "bubble_2d", "eos", "closed"])
def test_fused_diagnostics_export_equals_the_generic_one(monkeypatch, case):
    """FlowUpdateDiagnostics on the fused path (k_export_diag_fused: one lane per fluid node and component over the position-indexed
    rows + a fill of the solid nodes) against the generic k_export (TXG_EXPORT_GENERIC=1, read at the call) on the same device state,
    and against the oracle."""
    if case == "porous":
        cfg, walls, rho = cases.porous_3d(40, 24, 20, rmin=3.0, rmax=6.0)
    elif case == "bubble_2d":
        cfg, walls, rho = cases.bubble_2d(64)
    elif case == "closed":
        cfg, walls, rho = cases.porous_3d(24, rmin=3.0, rmax=6.0, periodic=(0, 1, 0))
    else:
        cfg, walls, rho = cases.porous_3d(32, rmin=4.0, rmax=8.0)
        cfg.use_nonideal_eos = 1
        for m in range(2):
            cfg.eos_type[m] = tc.EOS_SC
            cfg.eos_rho0[m] = 0.8 + 0.3 * m
    monkeypatch.delenv("TXG_EXPORT_GENERIC", raising=False)
    flow = gpu_util.make_flow(cfg, walls, rho)
    flow.step(15)
    fast = [a.copy() for a in flow.update_diagnostics()]
    assert flow.kernel_times()["k_export_diag_fused"][1] == 1
    monkeypatch.setenv("TXG_EXPORT_GENERIC", "1")
    slow = [a.copy() for a in flow.update_diagnostics()]
    assert flow.kernel_times()["k_export"][1] >= 1
    flow.close()
    o = cases.run_oracle(cfg, walls, rho, 15)
    for a, b, c, name in zip(fast, slow, o.diagnostics(), ("rhot", "prs", "velt")):
        assert gpu_util.rel_err(a, b) <= 1e-14, (name, gpu_util.rel_err(a, b))
        assert gpu_util.rel_err(a, c) <= TOL, name
        assert np.array_equal(a == 0, b == 0) or name == "prs"


def test_delta_norm_vs_oracle():
    cfg, walls, rho = cases.bubble_3d(16, hw=3)
    o = cases.run_oracle(cfg, walls, rho, 5)
    flow = gpu_util.make_flow(cfg, walls, rho)
    flow.step(5)
    assert flow.delta_norm() == 1e99 and o.delta_norm() == 1e99
    o.step(1)
    flow.step(1)
    a, b = flow.delta_norm(), o.delta_norm()
    assert abs(a - b) <= 1e-9 * b


def test_mass_conservation_large_porous():
    """Size-independent property at a size the oracle would not finish quickly."""
    cfg, walls, rho = cases.porous_3d(128)
    flow = gpu_util.make_flow(cfg, walls, rho)
    fluid = walls == 0
    m0 = gpu_util.mass(rho, fluid)
    flow.step(200)
    r = geo.owned(flow.get_arrays(u=False, forces=False)[0], cfg.stencil_size_rho, 3)
    m1 = gpu_util.mass(r, fluid)
    assert np.all(np.abs(m1 - m0) <= 1e-12 * np.abs(m0)), (m0, m1)
    assert np.all(r[~fluid] == 0)
    assert np.isfinite(r).all()


def test_output_files_match_reference_golden(tmp_path):
    """LBMOutput through the device path: fi001.dat of the shipped bubble_2D case, compared the way the
    reference's `make test` does (src/testing/check_solution.py: max |a - b| < eps = 1e-5) and at round-off."""
    from pathlib import Path

    from taxila_lbm_b200 import petsc_io

    cfg, walls, rho = cases.bubble_2d()
    flow = gpu_util.make_flow(cfg, walls, rho)
    flow.step(100)
    paths = flow.output_diagnostics(str(tmp_path) + "/", 1)
    names = sorted(Path(q).name for q in paths)
    assert names == ["fi001.dat", "prs001.dat", "rho001.dat", "rhot001.dat", "u001.dat"]
    got = petsc_io.read_vec(tmp_path / "fi001.dat")
    ref = petsc_io.read_vec(Path(__file__).resolve().parent / "golden" / "bubble_2D_fi001.dat")
    assert got.size == ref.size
    assert np.abs(got - ref).max() < 1e-5
    assert np.abs(got - ref).max() <= 1e-12
    # rho file = sum over directions of the fi file; rhot = sum over components (mm = 1)
    fi = got.reshape(128, 128, 9, 2)
    r = petsc_io.read_vec(tmp_path / "rho001.dat").reshape(128, 128, 2)
    assert np.abs(r - fi.sum(axis=2)).max() <= 1e-13
    rt = petsc_io.read_vec(tmp_path / "rhot001.dat").reshape(128, 128)
    assert np.abs(rt - r.sum(axis=2)).max() <= 1e-13
    flow.close()
