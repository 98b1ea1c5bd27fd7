"""The oracle's restatement of lbm_bc.F90 (external face BCs) and of the free-slip branch of
DistributionBouncebackD*.  The reference ships no golden vector for any of them ("parity unpinned"
rows of DESIGN.md), so the restatement is checked against what each node routine is built to
achieve -- the commented-out "recalc and affirm it worked" block of BCApplyNeumannNode
(lbm_bc.F90:1580-1590) -- and against structural facts of the step."""
import numpy as np
import pytest

import cases
from taxila_lbm_b200 import config as tc


def _face(a, boundary, D):
    """face plane of a natural-order array [z][y][x][...]"""
    axis = boundary // 2
    idx = 0 if boundary % 2 == 0 else -1
    sl = [slice(None)] * 3
    sl[2 - axis] = idx
    return a[tuple(sl)]


def _moments(o, cfg):
    lat = o.lattice()
    fi = o.fi()  # [z][y][x][Q][S]
    rho = fi.sum(axis=3)
    mom = np.einsum("zyxqs,qd->zyxds", fi, lat["ci"][:, :cfg.ndims].astype(float))
    return fi, rho, mom


def test_dirichlet_face_reaches_the_prescribed_density():
    cfg, walls, rho, bcs = cases.channel_2d(inlet=tc.BC_DIRICHLET, outlet=tc.BC_DIRICHLET)
    o = cases.run_oracle_bc(cfg, walls, rho, bcs, 25)
    fi, r, mom = _moments(o, cfg)
    for b in (tc.BOUNDARY_XM, tc.BOUNDARY_XP):
        fluid = _face(walls, b, 2) == 0
        want = bcs[b][..., 0, :]                        # pvals(m,1)
        got = _face(r, b, 2)[0]
        assert np.abs(got - want)[fluid[0]].max() < 1e-14
        # tangential momentum of every component is driven to zero (Q(p) = -momentum(p)/weightsum(p))
        assert np.abs(_face(mom, b, 2)[0][..., 1, :][fluid[0]]).max() < 1e-15
    # BCUpdateRho: the stored density of the face nodes is the sum of the corrected populations
    assert np.abs(o.rho() - r)[walls == 0].max() < 1e-15


def test_neumann_face_reaches_the_prescribed_momentum():
    cfg, walls, rho, bcs = cases.drainage_3d(inlet=tc.BC_NEUMANN, outlet=tc.BC_DIRICHLET)
    o = cases.run_oracle_bc(cfg, walls, rho, bcs, 12)
    fi, r, mom = _moments(o, cfg)
    F = o.forces()
    b = tc.BOUNDARY_ZM
    fluid = _face(walls, b, 3) == 0
    got = _face(mom, b, 3) + 0.5 * _face(F, b, 3)        # momentum + forces/2 (lbm_bc.F90:1588)
    assert np.abs(got - bcs[b])[fluid].max() < 1e-15
    b = tc.BOUNDARY_ZP
    fluid = _face(walls, b, 3) == 0
    assert np.abs(_face(r, b, 3) - bcs[b][..., 0, :])[fluid].max() < 1e-14


def test_velocity_face_reaches_the_prescribed_velocity():
    cfg, walls, rho, bcs = cases.drainage_3d(inlet=tc.BC_VELOCITY, outlet=tc.BC_DIRICHLET)
    o = cases.run_oracle_bc(cfg, walls, rho, bcs, 12)
    fi, r, mom = _moments(o, cfg)
    F = o.forces()
    b = tc.BOUNDARY_ZM
    fluid = _face(walls, b, 3) == 0
    with np.errstate(divide="ignore", invalid="ignore"):  # solid face nodes hold rho = 0; they are masked below
        u = (_face(mom, b, 3) + 0.5 * _face(F, b, 3)) / _face(r, b, 3)[..., None, :]
    want = bcs[b][..., :, 0]                              # uvals(1,:) for every component
    for m in range(2):
        assert np.abs(u[..., m] - want)[fluid].max() < 1e-13


@pytest.mark.parametrize("case", ["channel", "drainage"])
def test_prestream_has_no_effect_behind_the_ghost_wall(case):
    """BCPreStream (lbm_bc.F90:613-779) parks the incoming populations in the ghost layer so that the
    stream returns them; the ghost layer of every non-periodic face is WALL_GHOST (lbm_walls.F90:190-231),
    so the bounce-back sweep overwrites exactly those slots afterwards.  The device path therefore has no
    pre-stream step; this is the evidence that none is needed."""
    if case == "channel":
        cfg, walls, rho, bcs = cases.channel_2d(inlet=tc.BC_VELOCITY, outlet=tc.BC_DIRICHLET, walls_kind="noslip")
    else:
        cfg, walls, rho, bcs = cases.drainage_3d(N=16, NZ=20)
    a = cases.run_oracle_bc(cfg, walls, rho, bcs, 8, prestream=True)
    b = cases.run_oracle_bc(cfg, walls, rho, bcs, 8, prestream=False)
    assert np.array_equal(a.fi(), b.fi())
    assert np.isfinite(a.fi()).all()


def test_unflagged_nonperiodic_face_is_a_bounceback_wall():
    cfg, walls, rho, bcs = cases.channel_2d(inlet=tc.BC_NULL, outlet=tc.BC_NULL, walls_kind="noslip")
    o = cases.run_oracle_bc(cfg, walls, rho, {}, 0)
    m0 = o.rho()[walls == 0].sum(axis=0)
    o.step(30)
    m1 = o.rho()[walls == 0].sum(axis=0)
    assert np.abs(m1 - m0).max() / m0.max() < 1e-13


def test_freeslip_duct_conserves_mass_and_tangential_symmetry():
    """initialize_walls_nostick_duct: WALL_NORMAL_Y rows, x periodic.  Specular reflection conserves the
    mass of each component, and a y-uniform state driven along x stays y-uniform (a no-slip wall would
    build a Poiseuille profile instead)."""
    NX, NY = 16, 12
    c = tc.default_config(2, 1, NX, NY, 1)
    c.periodic[0], c.periodic[1] = 1, 0
    c.body_forces = 1
    c.gvt[0] = 1e-4
    tc.finalize_flags(c)
    walls = np.zeros((1, NY, NX))
    walls[0, 0, :] = walls[0, -1, :] = tc.WALL_NORMAL_Y
    rho = np.ones((1, NY, NX, 1))
    rho[walls != 0] = 0
    o = cases.run_oracle_bc(c, walls, rho, {}, 40)
    r, u = o.rho(), o.u()
    fluid = walls == 0
    assert abs(r[fluid].sum() - rho[fluid].sum()) < 1e-11
    ux = u[0, 1:-1, :, 0, 0]
    assert ux.min() > 30 * 1e-4                     # accelerating plug flow
    assert np.ptp(ux) < 1e-15                       # no shear at a free-slip wall
    # the same duct with no-slip walls does develop shear
    walls2 = np.where(walls != 0, 1.0, 0.0)
    o2 = cases.run_oracle_bc(c, walls2, rho, {}, 40)
    assert np.ptp(o2.u()[0, 1:-1, :, 0, 0]) > 1e-4


def test_reflecting_pairs_restated_literally():
    """BCApplyReflectingD3 (lbm_bc.F90:825-977): mirror pairs on every face except xm, whose test
    compares ci(n,X) with -ci(p,Z) (:849) and therefore only rewrites EASTDOWN, last from WESTDOWN."""
    cfg, walls, rho, bcs = cases.drainage_3d(N=8, NZ=8)
    import oracle

    o = oracle.Oracle(cfg)
    ci = o.lattice()["ci"]
    for b in range(1, 6):
        axis = b // 2
        pairs = o.reflecting_pairs(b)
        assert len(pairs) == 5
        for n, p in pairs:
            want = ci[n].copy()
            want[axis] = -want[axis]
            assert (ci[p] == want).all()
    pairs = o.reflecting_pairs(0)
    assert [tuple(ci[n]) for n, _ in pairs] == [(1, 0, -1)] * 3
    assert tuple(ci[pairs[-1][1]]) == (1, 0, -1) and tuple(ci[pairs[-2][1]]) == (-1, 0, -1)


def test_reflecting_face_keeps_a_stale_density():
    """BCUpdateRho (lbm_bc.F90:436-456) runs for Dirichlet / Neumann / velocity faces only: after BCApplyReflecting has
    rewritten the incoming populations of a reflecting face, dist%rho there is still the sum from before, and the
    next collision uses it.  (Why the device path refuses BC_REFLECTING instead of summing the populations.)"""
    cfg, walls, rho, bcs = cases.drainage_3d(N=12, NZ=12, inlet=tc.BC_DIRICHLET, outlet=tc.BC_DIRICHLET, x_bc=tc.BC_REFLECTING)
    o = cases.run_oracle_bc(cfg, walls, rho, bcs, 5)
    fi, r, mom = _moments(o, cfg)
    stored = o.rho()
    interior = (walls == 0)
    interior[:, :, 0] = interior[:, :, -1] = False
    assert np.abs(stored - r)[interior].max() < 1e-15
    xp = (walls == 0)[:, :, -1]
    assert np.abs(stored - r)[:, :, -1][xp].max() > 1e-6


def test_pressure_outlet_rederives_the_face_densities():
    """FlowUpdateBCPressureOutlet + FlowUpdateDensityFromPressure (lbm_flow.F90:1993-2263): the densities put on
    the outlet face give the prescribed pressure p = (rho_1 + rho_2)/3 + c_0 g_21 rho_1 rho_2 at the phase
    fraction of the node one step inside."""
    cfg, walls, rho, bcs = cases.channel_2d(inlet=tc.BC_VELOCITY, outlet=tc.BC_DIRICHLET, walls_kind="noslip")
    p_out = 0.31
    o = cases.run_oracle_bc(cfg, walls, rho, bcs, 20, outlets={tc.BOUNDARY_XP: p_out})
    fi, r, mom = _moments(o, cfg)
    fluid = walls[0, :, -1] == 0
    face = r[0, :, -1][fluid]                        # BCUpdateRho: what the Dirichlet correction reached
    g21 = cfg.gf[1][0]
    p = (face[:, 0] + face[:, 1]) / 3 + 6.0 * g21 * face[:, 0] * face[:, 1]
    assert np.abs(p - p_out).max() < 1e-13
    # and it differs from the constant-density outlet the face array started with
    assert np.abs(face - bcs[tc.BOUNDARY_XP][..., 0, :][fluid]).max() > 1e-3
