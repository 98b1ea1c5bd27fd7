"""Shared case builders for the parity tests (the configs of BASELINE.json / SURVEY.md 8d)."""
import numpy as np

import taxila_lbm_b200  # noqa: F401
from taxila_lbm_b200 import config as tc
from taxila_lbm_b200 import geometry as geo


from taxila_lbm_b200.workloads import bubble_2d, bubble_2d_hots, bubble_3d, porous_3d, porous_config  # noqa: E402,F401


def golden_bubble_2d():
    from pathlib import Path

    p = Path(__file__).resolve().parent / "golden" / "bubble_2D_fi001.dat"
    raw = np.fromfile(p, dtype=">f8", offset=8)
    hdr = np.fromfile(p, dtype=">i4", count=2)
    assert hdr[0] == 1211214 and hdr[1] == raw.size == 128 * 128 * 9 * 2
    return raw.astype(np.float64).reshape(1, 128, 128, 9, 2)


def run_oracle(cfg, walls, rho, steps, threads=1):
    import oracle

    o = oracle.Oracle(cfg, threads=threads)
    o.set_walls(walls)
    o.set_rho(rho)
    o.fi_init()
    o.update_moments()
    o.step(steps)
    return o


def bc_values(cfg, boundary, kind, rho=(1.0, 0.0), u=None, nz_local=None):
    """Constant face array of one boundary in the reference's layout ([t2][t1][ndims][S], what
    initialize_bcs_constant / FlowSetUpBCsD* fill): DIRICHLET densities in (0, m); NEUMANN momentum
    rho_m * u_d in (d, m); VELOCITY u_d in (d, 0)."""
    D, S = cfg.ndims, cfg.ncomponents
    NZ = (cfg.NZ if nz_local is None else nz_local) if D == 3 else 1
    n = {0: (NZ, cfg.NY), 1: (NZ, cfg.NX), 2: (cfg.NY, cfg.NX)}[boundary // 2]
    if D == 2:
        n = n[1:]
    v = np.zeros(n + (D, S))
    u = (0.0,) * D if u is None else u
    if kind == tc.BC_DIRICHLET:
        for m in range(S):
            v[..., 0, m] = rho[m]
    elif kind == tc.BC_NEUMANN:
        for m in range(S):
            for d in range(D):
                v[..., d, m] = rho[m] * u[d]
    elif kind == tc.BC_VELOCITY:
        for d in range(D):
            v[..., d, 0] = u[d]
    return v


def channel_2d(NX=48, NY=24, inlet=None, outlet=None, walls_kind="periodic", mrt=False, g=0.1, seed=5):
    """D2Q9 two-component flow along x between an inlet (xm) and an outlet (xp) face BC.
    walls_kind: 'periodic' (y periodic), 'noslip' (mineral rows at y = 0, NY-1), 'freeslip'
    (WALL_NORMAL_Y rows, the reference's initialize_walls_nostick_duct).  A few solid discs sit in the
    channel away from the side walls.  Returns (cfg, walls, rho, bcs) with bcs = {boundary: array}."""
    inlet = tc.BC_DIRICHLET if inlet is None else inlet
    outlet = tc.BC_DIRICHLET if outlet is None else outlet
    c = tc.default_config(2, 2, NX, NY, 1)
    c.periodic[0] = 0
    c.periodic[1] = 1 if walls_kind == "periodic" else 0
    c.relaxation_mode = tc.RELAXATION_MODE_MRT if mrt else tc.RELAXATION_MODE_SRT
    if mrt:
        for m in range(2):
            c.s_e[m], c.s_e2[m], c.s_q[m], c.s_nu[m] = 1.19, 1.4, 1.2, 0.9
    else:
        c.tau[0], c.tau[1] = 1.0, 0.9
    c.gf[0][1] = c.gf[1][0] = g
    c.gw[0][0], c.gw[0][1] = -0.03, 0.03
    c.bc_flags[tc.BOUNDARY_XM] = inlet
    c.bc_flags[tc.BOUNDARY_XP] = outlet
    tc.finalize_flags(c)
    walls = np.zeros((1, NY, NX))
    if walls_kind == "noslip":
        walls[0, 0, :] = walls[0, NY - 1, :] = 1.0
    elif walls_kind == "freeslip":
        walls[0, 0, :] = walls[0, NY - 1, :] = tc.WALL_NORMAL_Y
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:NY, 0:NX]
    for _ in range(3):
        cx, cy = rng.uniform(8, NX - 8), rng.uniform(7, NY - 7)
        walls[0][(xx - cx) ** 2 + (yy - cy) ** 2 <= 2.5 ** 2] = 1.0
    rho = np.zeros((1, NY, NX, 2))
    rho[..., 0] = np.where(xx < NX // 3, 0.95, 0.05)
    rho[..., 1] = 1.0 - rho[..., 0]
    rho[walls != 0] = 0.0
    bcs = {
        tc.BOUNDARY_XM: bc_values(c, tc.BOUNDARY_XM, inlet, rho=(0.99, 0.04), u=(0.02, 0.003)),
        tc.BOUNDARY_XP: bc_values(c, tc.BOUNDARY_XP, outlet, rho=(0.05, 0.93), u=(0.015, -0.002)),
    }
    return c, walls, rho, bcs


def drainage_3d(N=24, NZ=32, inlet=None, outlet=None, order=4, mrt=True, seed=11, x_bc=None):
    """D3Q19 two-component drainage along z through random spheres: zm inlet and zp outlet face BCs,
    x and y periodic (or, with x_bc, a face BC on xm/xp too -- edge nodes then see two BCs in sequence).
    Physics of C4 (MRT, 3 minerals, body force)."""
    inlet = tc.BC_NEUMANN if inlet is None else inlet
    outlet = tc.BC_DIRICHLET if outlet is None else outlet
    per = (0 if x_bc else 1, 1, 0)
    c, walls, rho = porous_3d(N, N, NZ, order=order, mrt=mrt, seed=seed, rmin=3.0, rmax=5.0, periodic=per, solid_fraction=0.3)
    c.bc_flags[tc.BOUNDARY_ZM] = inlet
    c.bc_flags[tc.BOUNDARY_ZP] = outlet
    if x_bc:
        c.bc_flags[tc.BOUNDARY_XM] = c.bc_flags[tc.BOUNDARY_XP] = x_bc
    bcs = {
        tc.BOUNDARY_ZM: bc_values(c, tc.BOUNDARY_ZM, inlet, rho=(0.98, 0.03), u=(0.001, -0.002, 0.01)),
        tc.BOUNDARY_ZP: bc_values(c, tc.BOUNDARY_ZP, outlet, rho=(0.04, 0.95), u=(0.0, 0.001, 0.008)),
    }
    if x_bc:
        bcs[tc.BOUNDARY_XM] = bc_values(c, tc.BOUNDARY_XM, x_bc, rho=(0.5, 0.5), u=(0.004, 0.0, 0.002))
        bcs[tc.BOUNDARY_XP] = bc_values(c, tc.BOUNDARY_XP, x_bc, rho=(0.45, 0.52), u=(-0.003, 0.001, 0.0))
    return c, walls, rho, bcs


def run_oracle_bc(cfg, walls, rho, bcs, steps, prestream=True, outlets=None):
    """outlets = {boundary: pressure}: faces flagged BC_PRESSURE_OUTLET (lbm_flow.F90:1170-1189)."""
    import oracle

    o = oracle.Oracle(cfg)
    o.set_walls(walls)
    for b, v in bcs.items():
        o.set_bc_values(b, v)
    for b, p in (outlets or {}).items():
        o.set_bc_pressure_outlet(b, p)
    o.set_prestream(prestream)
    o.set_rho(rho)
    o.fi_init()
    o.update_moments()
    o.step(steps)
    return o


def eos_pr_thermo_3d(N=32):
    """-flow_use_nonideal_eos with a Peng-Robinson component (self-attraction g_11 < 0, T = 0.95 T_c) and a
    Shan-Chen '94 "thermo" component psi = psi0 exp(-rho0 / rho) (lbm_eos.F90:231-349), smooth initial state,
    C4's walls, minerals and body force."""
    cfg, walls, rho = porous_3d(N, order=4, rmin=N / 8.0, rmax=N / 4.0)
    cfg.use_nonideal_eos = 1
    cfg.eos_type[0] = tc.EOS_PR
    cfg.gf[0][0] = -0.5
    cfg.gf[0][1] = cfg.gf[1][0] = 0.05
    tc.set_eos_pr(cfg, 0, reduced_T=0.95)
    cfg.eos_type[1] = tc.EOS_THERMO
    cfg.eos_psi0[1], cfg.eos_rho0[1] = 1.2, 0.4
    tc.finalize_flags(cfg)
    zz, yy, xx = np.mgrid[0:N, 0:N, 0:N]
    rho = np.zeros((N, N, N, 2))
    rho[..., 0] = 0.8 + 0.1 * np.sin(2 * np.pi * xx / N) * np.cos(2 * np.pi * zz / N)
    rho[..., 1] = 0.5 + 0.1 * np.cos(2 * np.pi * yy / N)
    rho[walls != 0] = 0
    return cfg, walls, rho
