"""GPU parity of SURVEY.md 8a-22 against the oracle: external face BCs (BCApplyDirichletToRho, BCApply,
BCUpdateRho of lbm_bc.F90 for reflecting / Dirichlet / Neumann / velocity faces) and free-slip walls
(WALL_NORMAL_X/Y/Z, lbm_distribution_function.F90:687-716), through the C ABI (txg_set_bc_values,
txg_config.bc_flags).  Tolerance: 1e-10 * max|field| on populations, density, velocity and forces, the
bar BASELINE.json states for the fields."""
import numpy as np
import pytest

import cases
import gpu_util
from taxila_lbm_b200 import capi
from taxila_lbm_b200 import config as tc
from taxila_lbm_b200 import geometry as geo

pytestmark = pytest.mark.gpu

TOL = 1e-10


def compare_bc(cfg, walls, rho, bcs, steps, tol=TOL, kernels=(), outlets=None):
    o = cases.run_oracle_bc(cfg, walls, rho, bcs, steps, outlets=outlets)
    flow = gpu_util.make_flow_bc(cfg, walls, rho, bcs, outlets=outlets)
    flow.step(steps)
    fi, r, u, F = gpu_util.fields(flow)
    fluid = np.asarray(walls).reshape(r.shape[:3]) == 0
    ofi = o.fi()
    assert np.isfinite(ofi[fluid]).all()
    errs = {
        "fi": gpu_util.rel_err(fi[fluid], ofi[fluid]),
        "rho": gpu_util.rel_err(r[fluid], o.rho()[fluid]),
        "u": gpu_util.rel_err(u[fluid], o.u()[fluid]),
        "forces": gpu_util.rel_err(F[fluid], o.forces()[fluid]) if np.abs(o.forces()).max() > 0 else 0.0,
    }
    for k, v in errs.items():
        assert v <= tol, (k, v, errs)
    assert np.all(fi[~fluid] == 0.0)
    launched = flow.kernel_times()
    for k in kernels:
        assert launched.get(k, (0, 0))[1] > 0, (k, launched)
    # FlowUpdateDiagnostics on the same state
    rhot, prs, velt = flow.update_diagnostics()
    ort, opr, ovt = o.diagnostics()
    assert gpu_util.rel_err(rhot[fluid], ort[fluid]) <= tol
    assert gpu_util.rel_err(prs[fluid], opr[fluid]) <= tol
    assert gpu_util.rel_err(velt[fluid], ovt[fluid]) <= tol
    flow.close()
    return errs


def test_pressure_driven_channel_2d():
    """bc_density / bc_pressure faces (BC_DIRICHLET) on xm and xp, y periodic, SRT."""
    compare_bc(*cases.channel_2d(inlet=tc.BC_DIRICHLET, outlet=tc.BC_DIRICHLET), steps=60,
               kernels=("k_bc_dirichlet_rho", "k_bc_apply", "k_step_fused", "k_forces_face", "k_collide_face"))


def test_velocity_inlet_noslip_channel_2d_mrt():
    compare_bc(*cases.channel_2d(inlet=tc.BC_VELOCITY, outlet=tc.BC_DIRICHLET, walls_kind="noslip", mrt=True), steps=60)


def test_flux_inlet_freeslip_duct_2d():
    """BC_NEUMANN inlet + the reference's nostick duct (WALL_NORMAL_Y rows): face BCs and mirrors together."""
    compare_bc(*cases.channel_2d(inlet=tc.BC_NEUMANN, outlet=tc.BC_DIRICHLET, walls_kind="freeslip"), steps=60,
               kernels=("k_specular_gather", "k_specular_scatter", "k_bc_apply"))


def test_pressure_outlet_2d_and_3d():
    """bc_pressure_outlet faces: the outlet densities follow the phase fraction arriving at the face
    (FlowUpdateBCPressureOutlet, lbm_flow.F90:1993-2263), re-derived on the device every step."""
    compare_bc(*cases.channel_2d(inlet=tc.BC_VELOCITY, outlet=tc.BC_DIRICHLET, walls_kind="noslip"), steps=60,
               outlets={tc.BOUNDARY_XP: 0.31}, kernels=("k_bc_pressure_outlet",))
    compare_bc(*cases.drainage_3d(inlet=tc.BC_NEUMANN, outlet=tc.BC_DIRICHLET), steps=30,
               outlets={tc.BOUNDARY_ZP: 0.30}, kernels=("k_bc_pressure_outlet",))


@pytest.mark.parametrize("order", [4, 8])
def test_drainage_3d_flux_inlet_pressure_outlet(order):
    """The drainage set-up the porous benchmark stands for: zm flux inlet, zp pressure outlet, MRT, minerals."""
    compare_bc(*cases.drainage_3d(inlet=tc.BC_NEUMANN, outlet=tc.BC_DIRICHLET, order=order), steps=40)


def test_drainage_3d_velocity_inlet_and_side_faces():
    """Velocity inlet; xm / xp are BC faces too, so edge nodes are corrected twice, in BCApply's order."""
    compare_bc(*cases.drainage_3d(inlet=tc.BC_VELOCITY, outlet=tc.BC_DIRICHLET, x_bc=tc.BC_NEUMANN), steps=30)


def test_reflecting_faces_3d():
    """BC_REFLECTING on xm / xp (BCApplyReflectingD3, lbm_bc.F90:825-977; the xm test compares ci(n, X) with -ci(p, Z) as
    written, :849) between a Dirichlet inlet and outlet on z: the fluid nodes of the reflecting faces collide with the
    density from before BCApply (BCUpdateRho skips them, :443-445), the edge nodes they share with the z faces do not."""
    compare_bc(*cases.drainage_3d(inlet=tc.BC_DIRICHLET, outlet=tc.BC_DIRICHLET, x_bc=tc.BC_REFLECTING), steps=40,
               kernels=("k_bc_reflect", "k_bc_apply", "k_step_fused", "k_collide_face"))


def test_face_bcs_on_the_split_kernels(monkeypatch):
    """TXG_SPLIT=1: the same face-BC steps on k_forces + k_collide over every node (the form of round 1; the default
    collides every node with the fused kernel and the face nodes a second time with their stored forces)."""
    monkeypatch.setenv("TXG_SPLIT", "1")
    compare_bc(*cases.drainage_3d(inlet=tc.BC_VELOCITY, outlet=tc.BC_DIRICHLET, x_bc=tc.BC_NEUMANN), steps=30,
               kernels=("k_forces", "k_collide", "k_bc_apply"))
    compare_bc(*cases.drainage_3d(inlet=tc.BC_DIRICHLET, outlet=tc.BC_DIRICHLET, x_bc=tc.BC_REFLECTING), steps=30,
               kernels=("k_forces", "k_collide", "k_bc_reflect"))
    compare_bc(*cases.channel_2d(inlet=tc.BC_NEUMANN, outlet=tc.BC_DIRICHLET, walls_kind="freeslip"), steps=40, kernels=("k_collide",))


def test_reflecting_faces_y_and_flux_inlet_3d():
    """reflecting ym / yp (the mirror rule) with a flux inlet and a pressure-like Dirichlet outlet, order-8 stencil"""
    c, walls, rho, bcs = cases.drainage_3d(order=8)
    c.periodic[1] = 0
    c.bc_flags[tc.BOUNDARY_YM] = c.bc_flags[tc.BOUNDARY_YP] = tc.BC_REFLECTING
    compare_bc(c, walls, rho, bcs, steps=30, kernels=("k_bc_reflect", "k_forces_tile"))


def test_reflecting_faces_2d():
    """BCApplyReflectingD2 on ym / yp (lbm_bc.F90:979-1073) of a pressure-driven channel; xm is refused (the reference
    reads ci(p, Z_DIRECTION) of a two-column array there, :1001)."""
    import taxila_lbm_b200 as tx

    c, walls, rho, bcs = cases.channel_2d(inlet=tc.BC_DIRICHLET, outlet=tc.BC_DIRICHLET)
    c.periodic[1] = 0
    c.bc_flags[tc.BOUNDARY_YM] = c.bc_flags[tc.BOUNDARY_YP] = tc.BC_REFLECTING
    compare_bc(c, walls, rho, bcs, steps=60, kernels=("k_bc_reflect",))
    c2, walls2, rho2, bcs2 = cases.channel_2d(inlet=tc.BC_DIRICHLET, outlet=tc.BC_DIRICHLET)
    c2.bc_flags[tc.BOUNDARY_XM] = tc.BC_REFLECTING
    with pytest.raises(capi.TaxilaGpuError) as e:
        tx.Flow(c2)
    assert e.value.code == 56 and "lbm_bc.F90:1001" in str(e.value)


def test_freeslip_duct_3d_fused_path():
    """No face BC: the fused step kernel + the free-slip slots.  z-normal mirrors, x and y periodic."""
    N, NZ = 20, 14
    c = tc.default_config(3, 2, N, N, NZ)
    c.periodic[0] = c.periodic[1] = 1
    c.periodic[2] = 0
    c.relaxation_mode = tc.RELAXATION_MODE_MRT
    for m in range(2):
        c.s_e[m], c.s_e2[m], c.s_q[m], c.s_pi[m], c.s_m[m] = 1.19, 1.4, 1.2, 1.4, 1.98
    c.gf[0][1] = c.gf[1][0] = 0.1
    c.body_forces = 1
    c.gvt[0] = 2e-5
    c.gw[0][0], c.gw[0][1] = -0.02, 0.02
    tc.finalize_flags(c)
    walls = np.zeros((NZ, N, N))
    walls[0] = walls[-1] = tc.WALL_NORMAL_Z
    zz, yy, xx = np.mgrid[0:NZ, 0:N, 0:N]
    walls[(xx - 9.5) ** 2 + (yy - 10) ** 2 + (zz - 6.5) ** 2 <= 9.0] = 1.0
    rho = np.zeros((NZ, N, N, 2))
    rho[..., 0] = np.where(xx < N // 2, 0.9, 0.1)
    rho[..., 1] = 1.0 - rho[..., 0]
    rho[walls != 0] = 0
    errs = compare_bc(c, walls, rho, {}, steps=50, kernels=("k_step_stage", "k_specular_scatter"))
    # specular reflection conserves the mass of each component
    flow = gpu_util.make_flow_bc(c, walls, rho, {})
    flow.step(50)
    r = gpu_util.fields(flow)[1]
    fluid = walls == 0
    m0, m1 = gpu_util.mass(rho, fluid), gpu_util.mass(r, fluid)
    assert np.all(np.abs(m1 - m0) <= 1e-12 * np.abs(m0)), (m0, m1, errs)
    flow.close()


def test_mirror_corner_is_refused():
    import taxila_lbm_b200 as tx

    c = tc.default_config(2, 1, 8, 8, 1)
    tc.finalize_flags(c)
    walls = np.zeros((1, 8, 8))
    walls[0, 0, :] = walls[0, -1, :] = tc.WALL_NORMAL_Y
    walls[0, :, 0] = walls[0, :, -1] = tc.WALL_NORMAL_X
    flow = tx.Flow(c)
    with pytest.raises(capi.TaxilaGpuError) as e:
        flow.walls_set_values(geo.ghosted(walls, 1, c.periodic, 2, wall_ghost=True))
    assert e.value.code == 56 and "free-slip" in str(e.value)
    flow.close()


def test_bc_on_a_periodic_axis_is_refused():
    import taxila_lbm_b200 as tx

    c, walls, rho = cases.bubble_2d(16, hw=3)
    c.bc_flags[tc.BOUNDARY_XM] = tc.BC_DIRICHLET
    with pytest.raises(capi.TaxilaGpuError) as e:
        tx.Flow(c)
    assert e.value.code == 62
