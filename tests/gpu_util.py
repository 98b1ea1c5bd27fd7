"""Drive the CUDA path through the C ABI the way lbm.F90 drives the reference (LBMInit2 + LBMRun2)."""
import numpy as np

import taxila_lbm_b200 as tx
from taxila_lbm_b200 import geometry as geo


def make_flow(cfg, walls, rho, device=0):
    """FlowSetUp, walls, initial state, FlowFiInit, FlowUpdateMoments on one GPU."""
    D = cfg.ndims
    R = cfg.stencil_size_rho
    flow = tx.Flow(cfg, device=device)
    walls_rg = geo.ghosted(walls, R, cfg.periodic, D, wall_ghost=True)
    flow.walls_set_values(walls_rg)
    rho_rg = geo.ghosted(rho, R, cfg.periodic, D)
    flow.initialize_state(rho_rg)
    flow.fi_init()
    flow.update_moments()
    return flow


def fields(flow):
    """(fi, rho, u, forces) in natural order, owned nodes only."""
    D = flow.D
    fi = geo.owned(flow.get_fi(), 1, D)
    rho, u, F = flow.get_arrays()
    return fi, geo.owned(rho, flow.R, D), geo.owned(u, 1, D), geo.owned(F, 1, D)


def rel_err(a, b):
    """max |a-b| / max |b| (the parity definition of SURVEY.md 8d)."""
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def mass(rho, fluid):
    """Per-component total mass over fluid nodes, accumulated in extended precision (a plain
    float64 axis-sum over ~1e6 values is itself only good to ~1e-11)."""
    return np.array([np.sum(rho[..., m][fluid], dtype=np.longdouble) for m in range(rho.shape[-1])], dtype=np.float64)


def make_flow_bc(cfg, walls, rho, bcs, device=0, outlets=None):
    """make_flow with face arrays (BCSetValues) uploaded before FlowFiInit; outlets = {boundary: pressure}."""
    D = cfg.ndims
    R = cfg.stencil_size_rho
    flow = tx.Flow(cfg, device=device)
    flow.walls_set_values(geo.ghosted(walls, R, cfg.periodic, D, wall_ghost=True))
    for b, v in bcs.items():
        flow.bc_set_values(b, v)
    for b, p in (outlets or {}).items():
        flow.bc_set_pressure_outlet(b, p)
    flow.initialize_state(geo.ghosted(rho, R, cfg.periodic, D))
    flow.fi_init()
    flow.update_moments()
    return flow
