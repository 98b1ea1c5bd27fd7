"""Structural checks of the CPU oracle -- properties the reference's algorithm has
whatever the geometry, used to guard the restatement where the reference ships no golden vector:
the OpenMP timing variant equals the literal serial sweep, per-component mass is conserved, solid nodes
hold nothing, the two lattices' results do not depend on how the box is oriented, and a planar specular
wall (codes 900-902) reflects without losing mass."""
import numpy as np

import cases
from taxila_lbm_b200 import config as tc
from taxila_lbm_b200 import geometry as geo


def small_porous(**kw):
    return cases.porous_3d(24, rmin=3.0, rmax=6.0, **kw)


def test_threaded_sweeps_equal_the_serial_sweep():
    """bench.py's CPU legs run the oracle with all host threads; that variant splits the bounce-back
    sweep into a push pass and a zeroing pass (oracle/taxila_oracle.c, bounceback()).  It must give the
    bits of the literal serial sweep."""
    cfg, walls, rho = small_porous()
    a = cases.run_oracle(cfg, walls, rho, 12, threads=1)
    b = cases.run_oracle(cfg, walls, rho, 12, threads=4)
    assert np.array_equal(a.fi(), b.fi())
    assert np.array_equal(a.rho(), b.rho())
    assert np.array_equal(a.u(), b.u())


def test_mass_conserved_and_solids_empty():
    cfg, walls, rho = small_porous()
    fluid = walls == 0
    o = cases.run_oracle(cfg, walls, rho, 60)
    fi = o.fi()
    assert np.all(fi[~fluid] == 0.0)
    for m in range(2):
        m0 = np.sum(rho[..., m][fluid], dtype=np.longdouble)
        m1 = np.sum(fi[..., m][fluid], dtype=np.longdouble)
        assert abs(m1 - m0) <= 1e-13 * abs(m0), (m, m0, m1)
    # the density field is the zeroth moment of the populations
    assert np.abs(o.rho()[fluid] - fi.sum(axis=3)[fluid]).max() <= 1e-15


def test_axis_permutation_invariance():
    """Swapping x and y of the whole problem (geometry, state, body force) swaps the result
    (the check behind src/testing/check_solution.py --rotate)."""
    cfg, walls, rho = small_porous()
    cfg.gvt[0], cfg.gvt[1], cfg.gvt[2] = 2e-5, 0.0, 1e-5
    a = cases.run_oracle(cfg, walls, rho, 30)
    cfg2 = cfg.copy()
    cfg2.gvt[0], cfg2.gvt[1] = cfg.gvt[1], cfg.gvt[0]
    walls2 = np.ascontiguousarray(walls.transpose(0, 2, 1))
    rho2 = np.ascontiguousarray(rho.transpose(0, 2, 1, 3))
    b = cases.run_oracle(cfg2, walls2, rho2, 30)
    ra, rb = a.rho(), b.rho().transpose(0, 2, 1, 3)
    assert np.abs(ra - rb).max() <= 1e-13 * np.abs(ra).max()
    ua, ub = a.u(), b.u().transpose(0, 2, 1, 3, 4)  # (z, y, x, d, m): swap the x and y velocity components too
    ub = ub[:, :, :, [1, 0, 2], :]
    assert np.abs(ua - ub).max() <= 1e-12 * max(np.abs(ua).max(), 1e-30)


def test_planar_specular_walls_conserve_mass():
    """A slit between two WALL_NORMAL_Z planes (code 902), periodic in x and y, driven along x: the
    specular sweep (lbm_distribution_function.F90:687-716) loses no mass, and -- unlike bounce-back
    walls -- does not brake the flow (free slip)."""
    def slit(code):
        cfg = tc.default_config(3, 1, 12, 10, 9)
        cfg.periodic[0] = cfg.periodic[1] = 1
        cfg.body_forces = 1
        cfg.gvt[0] = 1e-4
        tc.finalize_flags(cfg)
        walls = np.zeros((9, 10, 12))
        walls[0] = walls[-1] = code
        rho = np.ones((9, 10, 12, 1))
        rho[walls != 0] = 0.0
        o = cases.run_oracle(cfg, walls, rho, 200)
        fluid = walls == 0
        return o, fluid, rho

    o, fluid, rho = slit(tc.WALL_NORMAL_Z)
    m0 = np.sum(rho[fluid], dtype=np.longdouble)
    m1 = np.sum(o.rho()[fluid], dtype=np.longdouble)
    assert abs(m1 - m0) <= 1e-12 * m0
    ux_slip = o.u()[..., 0, 0][fluid]
    ob, fb, _ = slit(tc.WALL_NONREACTIVE)
    ux_noslip = ob.u()[..., 0, 0][fb]
    # free slip: plug flow accelerating uniformly (200 steps x 1e-4); no slip: slower, with a profile
    assert np.ptp(ux_slip) <= 1e-10
    assert abs(ux_slip.mean() - 200 * 1e-4) <= 2e-4
    assert ux_noslip.mean() < 0.8 * ux_slip.mean() and np.ptp(ux_noslip) > 1e-3


def test_eos_variants_against_their_formulas():
    """EOSApply_SC / _Thermo / _PR (lbm_eos.F90:183-349) as the pressure term of FlowUpdateDiagnostics
    shows them: prs = rhot/3 + (c_0/2) sum_m psi_m sum_m' g_mm' psi_m' (lbm_flow.F90:654-758)."""
    import cases

    cfg, walls, rho = cases.eos_pr_thermo_3d(12)
    o = cases.run_oracle(cfg, walls, rho, 0)
    assert o.eos_bad() == 0
    r = o.rho()
    fluid = walls == 0
    f32 = lambda v: float(np.float32(v))  # noqa: E731
    om, T, Tc = cfg.eos_pr_omega[0], cfg.eos_pr_T[0], cfg.eos_pr_Tc[0]
    assert abs(Tc - (2 / 49) / (2 / 21) * f32(0.0778) / f32(0.45724)) < 1e-17 and abs(T - 0.95 * Tc) < 1e-17
    alpha = (1 + (f32(0.37464) + f32(1.54226) * om - f32(0.26992) * om ** 2) * (1 - np.sqrt(T / Tc))) ** 2
    a, b, g = cfg.eos_pr_a[0], cfg.eos_pr_b[0], cfg.gf[0][0]
    r0, r1 = r[..., 0][fluid], r[..., 1][fluid]
    psi0 = np.sqrt(2 * (r0 * T / (1 - b * r0) - a * alpha * r0 ** 2 / (1 + 2 * b * r0 - (b * r0) ** 2) - r0 / 3) / (6 * g))
    psi1 = cfg.eos_psi0[1] * np.exp(-cfg.eos_rho0[1] / r1)
    want = (r0 + r1) / 3 + 3 * (psi0 * (g * psi0 + cfg.gf[0][1] * psi1) + psi1 * (cfg.gf[1][0] * psi0 + cfg.gf[1][1] * psi1))
    prs = o.diagnostics()[1][fluid]
    assert np.abs(prs - want).max() <= 1e-13 * np.abs(want).max()
