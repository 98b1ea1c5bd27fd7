"""Workload builders of the configurations BASELINE.json names (C1 bubble_2D, C2 bubble_2D_hots, C3 bubble_3D, C4 porous
drainage): config + walls + initial densities in natural order.  Host set-up only; used by bench.py, the smoke test
and the parity tests (tests/cases.py re-exports them)."""
import numpy as np

from . import config as tc
from . import geometry as geo


def bubble_2d(N=128, mrt=False, order=4, g=0.1, rho_in=(0.03, 0.97), rho_out=(0.97, 0.03), hw=26):
    """C1 tests/bubble_2D (as shipped) and variations."""
    c = tc.default_config(2, 2, N, N, 1)
    c.periodic[0] = c.periodic[1] = 1
    c.relaxation_mode = tc.RELAXATION_MODE_MRT if mrt else tc.RELAXATION_MODE_SRT
    c.isotropy_order = order
    c.gf[0][1] = c.gf[1][0] = g
    tc.finalize_flags(c)
    walls = np.zeros((1, N, N))
    rho = geo.bubble_rho(c, rho_in, rho_out, hw)
    return c, walls, rho


def bubble_2d_hots(N=128):
    """C2 tests/bubble_2D_hots/input_data: MRT, derivative order 10, gvt = (0,0)."""
    c, walls, rho = bubble_2d(N, mrt=True, order=10, g=0.01666666666, rho_in=(0.01, 0.99), rho_out=(0.99, 0.01))
    rates = dict(s_c=(1.0, 1.0), s_e=(0.1, 1.8), s_e2=(0.2, 1.8), s_q=(0.625, 1.8), s_nu=(1.0, 1.0))
    for k, v in rates.items():
        for m in range(2):
            getattr(c, k)[m] = v[m]
    c.body_forces = 1
    tc.finalize_flags(c)
    return c, walls, rho


def bubble_3d(N=128, NZ=None, mrt=False, order=4, hw=10):
    """C3 tests/bubble_3D."""
    NZ = N if NZ is None else NZ
    c = tc.default_config(3, 2, N, N, NZ)
    c.periodic[0] = c.periodic[1] = c.periodic[2] = 1
    c.relaxation_mode = tc.RELAXATION_MODE_MRT if mrt else tc.RELAXATION_MODE_SRT
    c.isotropy_order = order
    c.gf[0][1] = c.gf[1][0] = 0.1
    tc.finalize_flags(c)
    walls = np.zeros((NZ, N, N))
    rho = geo.bubble_rho(c, (0.03, 0.97), (0.97, 0.03), hw)
    return c, walls, rho


def porous_3d(NX=64, NY=None, NZ=None, order=4, mrt=True, seed=20260, rmin=10.0, rmax=22.0, periodic=(1, 1, 1),
              solid_fraction=0.55, walls=None):
    """C4 porous drainage: MRT rates s_c=1 s_nu=1 s_e=1.19 s_e2=1.4 s_q=1.2
    s_pi=1.4 s_m=1.98, g12=g21=0.1, 3 minerals gw(k) = (-0.02k, +0.02k), gvt=(0,0,1e-5),
    random overlapping spheres, flushing IC along z."""
    NY = NX if NY is None else NY
    NZ = NX if NZ is None else NZ
    c = porous_config(NX, NY, NZ, order=order, mrt=mrt, periodic=periodic)
    if walls is None:
        walls = geo.porous_spheres(NX, NY, NZ, seed=seed, rmin=rmin, rmax=rmax, solid_fraction=solid_fraction,
                                   nminerals=3, periodic=all(periodic))
    rho = geo.flushing_rho(c, walls, (0.97, 0.03), (0.03, 0.97), "z", 10)
    return c, walls, rho


def porous_config(NX, NY, NZ, order=4, mrt=True, periodic=(1, 1, 1)):
    """The parameters of C4 without the geometry."""
    c = tc.default_config(3, 2, NX, NY, NZ)
    for d in range(3):
        c.periodic[d] = periodic[d]
    c.relaxation_mode = tc.RELAXATION_MODE_MRT if mrt else tc.RELAXATION_MODE_SRT
    c.isotropy_order = order
    for m in range(2):
        c.s_c[m], c.s_nu[m], c.s_e[m], c.s_e2[m] = 1.0, 1.0, 1.19, 1.4
        c.s_q[m], c.s_pi[m], c.s_m[m] = 1.2, 1.4, 1.98
    c.gf[0][1] = c.gf[1][0] = 0.1
    c.nminerals = 3
    for k in range(3):
        c.gw[k][0] = -0.02 * (k + 1)
        c.gw[k][1] = +0.02 * (k + 1)
    c.body_forces = 1
    c.gvt[0], c.gvt[1], c.gvt[2] = 0.0, 0.0, 1.0e-5
    tc.finalize_flags(c)
    return c
