"""The oracle against a second, independently written model (tests/textbook_lbm.py) on the pieces that no
reference vector pins (SURVEY.md 8c "parity unpinned"): D3Q19 MRT with general rates, the order-8 / order-10
force stencils with their line-of-sight rules, bounce-back, mineral (fluid-solid) and body forces.
Two restatements that share no code agree to accumulated round-off; the textbook model itself is held to
the reference's golden fi001.dat of tests/bubble_2D."""
import numpy as np
import pytest

import cases
import textbook_lbm as tb
from taxila_lbm_b200 import config as tc
from taxila_lbm_b200 import geometry as geo

TOL = 1e-12  # relative to max |field|; the two models sum in different orders (measured: <= 1e-13)


def rel(a, b):
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def compare(cfg, walls, rho, steps, tol=TOL):
    o = cases.run_oracle(cfg, walls, rho, steps)
    t = tb.from_config(cfg, walls, rho)
    t.step(steps)
    fluid = np.asarray(walls).reshape(o.rho().shape[:3]) == 0
    errs = {"fi": rel(t.fi_natural()[fluid], o.fi()[fluid]), "rho": rel(t.rho_natural()[fluid], o.rho()[fluid]),
            "u": rel(t.u_natural()[fluid], o.u()[fluid][..., 0])}
    assert all(np.array_equal(o.u()[..., 0], o.u()[..., m]) for m in range(cfg.ncomponents))  # one common velocity
    if np.abs(o.forces()).max() > 0:
        errs["forces"] = rel(t.forces_natural()[fluid], o.forces()[fluid])
    assert all(v <= tol for v in errs.values()), errs
    assert np.all(o.fi()[~fluid] == 0.0)
    o.close()
    return errs


def test_textbook_model_reproduces_the_reference_golden():
    """tests/bubble_2D/reference_solution/fi001.dat (100 steps) from the textbook model alone."""
    cfg, walls, rho = cases.bubble_2d()
    t = tb.from_config(cfg, walls, rho)
    t.step(100)
    assert np.abs(t.fi_natural() - cases.golden_bubble_2d()).max() <= 1e-12


def test_moment_matrices_match_the_polynomials():
    """The oracle's moment rows (lbm_discretization_d3q19.F90:177-197, _d2q9.F90:108-118) are the
    d'Humieres / Lallemand-Luo polynomials: same row norms, and the collision operator M^-1 S M built from the
    polynomials equals the oracle's (checked through the runs below); here the norms the reference hard-codes."""
    assert np.allclose((tb.Lattice(3).M ** 2).sum(axis=1),
                       [19, 2394, 252, 10, 40, 10, 40, 10, 40, 36, 72, 12, 24, 4, 4, 4, 8, 8, 8])
    assert np.allclose((tb.Lattice(2).M ** 2).sum(axis=1), [9, 36, 36, 6, 12, 6, 12, 4, 4])
    for D in (2, 3):
        M = tb.Lattice(D).M
        G = M @ M.T
        assert np.allclose(G, np.diag(np.diag(G)))


def test_d2q9_mrt_general_rates_order10():
    """C2 (tests/bubble_2D_hots/input_data): MRT rates all different, derivative order 10 (36 offsets, radius 3)."""
    compare(*cases.bubble_2d_hots(48), steps=40)


def test_d2q9_order8_with_walls_and_minerals():
    """order-8 stencil (24 offsets) next to walls: every line-of-sight rule of the 2-D blocks is exercised."""
    cfg, walls, rho = cases.bubble_2d(40, mrt=True, order=8)
    cfg.stencil_size_rho = 2
    cfg.nminerals = 2
    for k in range(2):
        cfg.gw[k][0], cfg.gw[k][1] = -0.03 * (k + 1), 0.02 * (k + 1)
    for m in range(2):
        cfg.s_e[m], cfg.s_e2[m], cfg.s_q[m], cfg.s_nu[m] = 1.3, 0.9, 1.7, 1.0 / (0.8 + 0.3 * m)
    cfg.body_forces = 1
    cfg.gvt[0], cfg.gvt[1] = 2e-5, -1e-5
    tc.finalize_flags(cfg)
    rng = np.random.default_rng(3)
    walls = np.zeros((1, 40, 40))
    solid = rng.random((40, 40)) < 0.22
    walls[0][solid] = 1 + (rng.integers(0, 2, size=(40, 40))[solid])
    rho = rho.copy()
    rho[walls != 0] = 0.0
    compare(cfg, walls, rho, steps=30)


@pytest.mark.parametrize("order", [4, 8])
def test_d3q19_porous_mrt_minerals_body_force(order):
    """The C4 recipe (MRT s_e=1.19 s_e2=1.4 s_q=1.2 s_pi=1.4 s_m=1.98, 3 minerals, gvt, bounce-back) at 24^3."""
    cfg, walls, rho = cases.porous_3d(24, order=order, rmin=3.0, rmax=6.0)
    if order == 8:
        cfg.stencil_size_rho = 2
    compare(cfg, walls, rho, steps=25)


def test_d3q19_random_solids_order8_every_sight_rule():
    """Random 30 % solids: isolated voxels put a wall on every kind of line of sight of the 92 offsets."""
    cfg, walls, rho = cases.porous_3d(16, order=8, rmin=3.0, rmax=6.0)
    cfg.stencil_size_rho = 2
    rng = np.random.default_rng(8)
    walls = np.where(rng.random((16, 16, 16)) < 0.3, 1.0 + rng.integers(0, 3, size=(16, 16, 16)), 0.0)
    rho = geo.flushing_rho(cfg, walls, (0.9, 0.1), (0.2, 0.8), "z", 5)
    compare(cfg, walls, rho, steps=15)


def test_d3q19_srt_unequal_masses():
    """mm /= 1: d_k = 1 - 2 / (3 mm) in the equilibrium, mm-weighted common velocity and body force."""
    cfg, walls, rho = cases.porous_3d(16, mrt=False, rmin=3.0, rmax=5.0)
    cfg.mm[0], cfg.mm[1] = 1.0, 1.6
    cfg.tau[0], cfg.tau[1] = 0.9, 1.3
    compare(cfg, walls, rho, steps=25)


def test_nonperiodic_box_is_a_box_inside_a_ghost_wall():
    """A non-periodic face with no BC is the WALL_GHOST = 999 layer (lbm_walls.F90:190-231): bounce-back off it, no
    mineral force from it, no fluid-fluid stencil through it.  The textbook model runs the same box padded with a ring of
    999 nodes (periodic wrap falls inside the ring), orders 4 and 8."""
    for order in (4, 8):
        cfg, walls, rho = cases.porous_3d(14, order=order, rmin=3.0, rmax=5.0, periodic=(0, 0, 0))
        cfg.stencil_size_rho = 2 if order == 8 else 1
        o = cases.run_oracle(cfg, walls, rho, 20)
        pad = 3
        wp = np.pad(walls, pad, constant_values=999.0)
        rp = np.pad(rho, [(pad, pad)] * 3 + [(0, 0)])
        pc = cfg.copy()
        pc.NX = pc.NY = pc.NZ = 14 + 2 * pad
        t = tb.from_config(pc, wp, rp)
        t.step(20)
        inner = (slice(pad, -pad),) * 3
        fluid = walls == 0
        assert rel(t.fi_natural()[inner][fluid], o.fi()[fluid]) <= TOL
        assert rel(t.forces_natural()[inner][fluid], o.forces()[fluid]) <= TOL
        assert rel(t.u_natural()[inner][fluid], o.u()[fluid][..., 0]) <= TOL
        o.close()


def test_nonideal_eos_kinds_in_the_force():
    """-flow_use_nonideal_eos: psi(rho) replaces rho in the fluid-fluid gradient and as its prefactor only (EOSApply from
    FlowCalcForces, lbm_flow.F90:795-801); Peng-Robinson + '94 thermo, and Shan-Chen '93 on both components."""
    compare(*cases.eos_pr_thermo_3d(16), steps=25)
    cfg, walls, rho = cases.porous_3d(16, rmin=3.0, rmax=5.0)
    cfg.use_nonideal_eos = 1
    for m in range(2):
        cfg.eos_type[m] = tc.EOS_SC
        cfg.eos_rho0[m] = 0.8 + 0.3 * m
    compare(cfg, walls, rho, steps=25)


def test_freeslip_slit_two_components():
    """WALL_NORMAL_Z planes (code 902, lbm_distribution_function.F90:687-716) as textbook specular reflection: two
    components with a density pattern, Shan-Chen force next to the walls, body force along x and y, MRT."""
    cfg = tc.default_config(3, 2, 14, 12, 10)
    for d in range(3):
        cfg.periodic[d] = 1
    cfg.relaxation_mode = tc.RELAXATION_MODE_MRT
    for m in range(2):
        cfg.s_e[m], cfg.s_e2[m], cfg.s_q[m], cfg.s_pi[m], cfg.s_m[m] = 1.19, 1.4, 1.2, 1.4, 1.98
    cfg.gf[0][1] = cfg.gf[1][0] = 0.1
    cfg.body_forces = 1
    cfg.gvt[0], cfg.gvt[1] = 1e-4, -5e-5
    tc.finalize_flags(cfg)
    walls = np.zeros((10, 12, 14))
    walls[0] = walls[-1] = tc.WALL_NORMAL_Z
    zz, yy, xx = np.mgrid[0:10, 0:12, 0:14]
    rho = np.zeros((10, 12, 14, 2))
    rho[..., 0] = 0.6 + 0.3 * np.sin(2 * np.pi * xx / 14) * np.cos(2 * np.pi * yy / 12) + 0.02 * zz
    rho[..., 1] = 1.0 - rho[..., 0]
    rho[walls != 0] = 0.0
    compare(cfg, walls, rho, steps=40)


def test_freeslip_duct_2d():
    """initialize_walls_nostick_duct in 2-D: WALL_NORMAL_Y rows (code 901), x periodic, SRT, order-8 stencil next to the rows."""
    cfg, walls, rho = cases.bubble_2d(32, order=8, hw=6)
    cfg.stencil_size_rho = 2
    cfg.body_forces = 1
    cfg.gvt[0] = 1e-4
    tc.finalize_flags(cfg)
    walls = np.zeros((1, 32, 32))
    walls[0, 0] = walls[0, -1] = tc.WALL_NORMAL_Y
    rho = rho.copy()
    rho[walls != 0] = 0.0
    compare(cfg, walls, rho, steps=40)


def _padded_with_faces(cfg, walls, rho, bcs, pad=3):
    """Textbook model of a box whose non-periodic axes carry face BCs: the box padded with a 999 ring along those axes (the
    reference's ghost layer), the faces as masks with their (constant) values, in BCApply's order."""
    D = cfg.ndims
    S = cfg.ncomponents
    w = walls if D == 3 else walls[0]
    r = rho if D == 3 else rho[0]
    nonper = [d for d in range(D) if not cfg.periodic[d]]          # d = 0: x
    padw = [(pad, pad) if (D - 1 - ax) in nonper else (0, 0) for ax in range(D)]  # array axes run z, y, x
    wp = np.pad(w, padw, constant_values=999.0)
    rp = np.pad(r, padw + [(0, 0)])
    pc = cfg.copy()
    for d in nonper:
        setattr(pc, ("NX", "NY", "NZ")[d], getattr(cfg, ("NX", "NY", "NZ")[d]) + 2 * pad)
    t = tb.from_config(pc, wp if D == 3 else wp[None], rp if D == 3 else rp[None])
    kinds = {tc.BC_DIRICHLET: "dirichlet", tc.BC_NEUMANN: "neumann", tc.BC_VELOCITY: "velocity"}
    faces = []
    for b in sorted(bcs):
        d, side = b // 2, b % 2
        ax = D - 1 - d
        mask = np.zeros(wp.shape, dtype=bool)
        idx = [slice(pad, -pad) if (D - 1 - a) in nonper else slice(None) for a in range(D)]
        idx[ax] = pad if side == 0 else wp.shape[ax] - pad - 1
        mask[tuple(idx)] = True
        vals = np.asarray(bcs[b]).reshape(-1, D, S)
        assert np.all(vals == vals[0])  # constant faces
        faces.append((kinds[cfg.bc_flags[b]], d, side, mask, vals[0]))
    first = {}
    for i, f in enumerate(faces):  # each BC type once, in the order of its first face; faces xm .. zp inside a type
        first.setdefault(f[0], i)
    faces.sort(key=lambda f: first[f[0]])
    inner = tuple(slice(pad, -pad) if (D - 1 - a) in nonper else slice(None) for a in range(D))
    return t, faces, inner


def _compare_faces(case, steps, outlets=None):
    cfg, walls, rho, bcs = case
    o = cases.run_oracle_bc(cfg, walls, rho, bcs, steps, outlets=outlets)
    t, faces, inner = _padded_with_faces(cfg, walls, rho, bcs)
    t.p["faces"] = faces
    if outlets:  # {boundary: pressure} -> {index in faces: pressure}
        t.p["outlets"] = {i: outlets[2 * f[1] + f[2]] for i, f in enumerate(faces) if (2 * f[1] + f[2]) in outlets}
        assert len(t.p["outlets"]) == len(outlets)
    t.step(steps)
    fluid = np.asarray(walls).reshape(o.rho().shape[:3]) == 0
    sel = (slice(None),) + inner if cfg.ndims == 2 else inner
    errs = {"fi": rel(t.fi_natural()[sel][fluid], o.fi()[fluid]), "rho": rel(t.rho_natural()[sel][fluid], o.rho()[fluid]),
            "u": rel(t.u_natural()[sel][fluid], o.u()[fluid][..., 0]), "forces": rel(t.forces_natural()[sel][fluid], o.forces()[fluid])}
    assert all(v <= TOL for v in errs.values()), errs
    o.close()


def test_dirichlet_faces_2d_and_3d():
    """bc_density / bc_pressure faces (BCApplyDirichletToRho, BCApplyDirichletNode, BCUpdateRho of lbm_bc.F90 in the order of
    FlowApplyBCs, lbm_flow.F90:1958-1991): the pressure-driven 2-D channel and the 3-D drainage box with density faces."""
    _compare_faces(cases.channel_2d(inlet=tc.BC_DIRICHLET, outlet=tc.BC_DIRICHLET), 40)
    _compare_faces(cases.drainage_3d(N=16, NZ=20, inlet=tc.BC_DIRICHLET, outlet=tc.BC_DIRICHLET), 25)


def test_flux_and_velocity_faces():
    """bc_flux (BCApplyNeumannNode) and bc_velocity (BCApplyVelocityNode) inlets with a density outlet; no-slip side walls
    in 2-D; and the 3-D box with Neumann faces on xm / xp as well, whose edge nodes are corrected by two faces in turn."""
    _compare_faces(cases.channel_2d(inlet=tc.BC_NEUMANN, outlet=tc.BC_DIRICHLET, walls_kind="noslip"), 40)
    _compare_faces(cases.channel_2d(inlet=tc.BC_VELOCITY, outlet=tc.BC_DIRICHLET, walls_kind="noslip", mrt=True), 40)
    _compare_faces(cases.drainage_3d(N=16, NZ=20, inlet=tc.BC_NEUMANN, outlet=tc.BC_DIRICHLET), 25)
    _compare_faces(cases.drainage_3d(N=16, NZ=20, inlet=tc.BC_VELOCITY, outlet=tc.BC_DIRICHLET, x_bc=tc.BC_NEUMANN), 25)


def test_pressure_outlet_faces():
    """bc_pressure_outlet (FlowUpdateBCPressureOutlet + FlowUpdateDensityFromPressure, lbm_flow.F90:1993-2263): the face
    densities are re-derived every step from the phase fraction one node inside."""
    _compare_faces(cases.channel_2d(inlet=tc.BC_VELOCITY, outlet=tc.BC_DIRICHLET, walls_kind="noslip"), 40,
                   outlets={tc.BOUNDARY_XP: 0.31})
    _compare_faces(cases.drainage_3d(N=16, NZ=20, inlet=tc.BC_NEUMANN, outlet=tc.BC_DIRICHLET), 25, outlets={tc.BOUNDARY_ZP: 0.30})


def test_three_components_and_one_component():
    """S = 3 (3 x 3 coupling matrix, unequal masses, two minerals) and S = 1 (no fluid-fluid force, SRT)."""
    cfg = tc.default_config(3, 3, 14, 14, 14)
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
    walls = geo.porous_spheres(14, 14, 14, seed=11, rmin=2.0, rmax=4.0, solid_fraction=0.4, nminerals=2)
    rho = 0.2 + 0.6 * np.random.default_rng(5).random((14, 14, 14, 3))
    rho[walls != 0] = 0.0
    compare(cfg, walls, rho, steps=25)

    cfg = tc.default_config(3, 1, 12, 12, 12)
    for d in range(3):
        cfg.periodic[d] = 1
    cfg.tau[0] = 0.8
    cfg.body_forces = 1
    cfg.gvt[0] = 1e-4
    tc.finalize_flags(cfg)
    walls = geo.porous_spheres(12, 12, 12, seed=3, rmin=2.0, rmax=3.0, solid_fraction=0.3, nminerals=1)
    rho = np.ones((12, 12, 12, 1))
    rho[walls != 0] = 0.0
    compare(cfg, walls, rho, steps=25)
