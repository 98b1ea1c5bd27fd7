"""Worker of the multi-rank tests: launched by torch.distributed.run, one rank per GPU.

    python -m torch.distributed.run --nproc-per-node N tests/mg_worker.py --case NAME --steps K --out FILE

Every rank cuts its z-slab from the same global case, drives the CUDA path through the C ABI
exactly like the single-rank tests, and rank 0 assembles the owned fields and compares them
with (a) the CPU oracle on the undecomposed box and (b) optionally a single-rank GPU run,
which must agree BIT FOR BIT: the arithmetic per node does not depend on the decomposition.
"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import numpy as np  # noqa: E402


def build(name):
    import cases

    if name == "porous_periodic":
        return cases.porous_3d(32, NZ=48, rmin=4.0, rmax=8.0)
    if name == "porous_iso8":
        return cases.porous_3d(28, NZ=40, order=8, rmin=4.0, rmax=7.0)
    if name == "porous_closed_box":
        return cases.porous_3d(32, NZ=37, mrt=False, rmin=4.0, rmax=8.0, periodic=(0, 0, 0))
    if name == "bubble_srt":
        return cases.bubble_3d(24, NZ=30, hw=5)
    if name == "thin_slabs":  # slabs thinner than the boundary/interior split needs
        return cases.porous_3d(24, NZ=10, rmin=2.0, rmax=3.0)
    if name == "freeslip_duct":  # y-normal mirrors (901), flow along periodic z: reflections cross the slab faces
        from taxila_lbm_b200 import config as tc

        N, NZ = 20, 24
        c, walls, rho = cases.bubble_3d(N, NZ=NZ, mrt=True, hw=4)
        c.periodic[1] = 0
        for m in range(2):
            c.s_e[m], c.s_e2[m], c.s_q[m], c.s_pi[m], c.s_m[m] = 1.19, 1.4, 1.2, 1.4, 1.98
        c.body_forces = 1
        c.gvt[0], c.gvt[2] = 1e-5, 3e-5
        c.gw[0][0], c.gw[0][1] = -0.02, 0.02
        tc.finalize_flags(c)
        walls = np.zeros((NZ, N, N))
        walls[:, 0, :] = walls[:, -1, :] = tc.WALL_NORMAL_Y
        zz, yy, xx = np.mgrid[0:NZ, 0:N, 0:N]
        walls[(xx - 9.5) ** 2 + (yy - 10) ** 2 + (zz - 12.2) ** 2 <= 10.0] = 1.0  # an obstacle astride the slab face
        rho = np.asarray(rho).copy()
        rho[walls != 0] = 0.0
        return c, walls, rho
    if name == "drainage_bc":  # zm flux inlet, zp pressure outlet, Neumann faces on xm/xp: face BCs on every slab
        from taxila_lbm_b200 import config as tc

        return cases.drainage_3d(N=24, NZ=34, x_bc=tc.BC_NEUMANN)
    raise SystemExit("unknown case " + name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", required=True)
    ap.add_argument("--single", action="store_true", help="also compare with a 1-rank GPU run on rank 0")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import cases
    import gpu_util
    import taxila_lbm_b200 as tx
    from taxila_lbm_b200 import geometry as geo
    from taxila_lbm_b200 import slab

    built = build(args.case)
    cfg, walls, rho = built[:3]
    bcs = built[3] if len(built) > 3 else {}
    c, walls_rg, rho_rg = slab.local_arrays(cfg, walls, rho, world, rank)
    ids = [tx.Flow.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    flow = tx.Flow(c, device=local_rank, nccl_id=ids[0])
    flow.walls_set_values(walls_rg)
    for b, v in slab.local_bc_values(cfg, bcs, world, rank).items():
        flow.bc_set_values(b, v)
    flow.initialize_state(rho_rg)
    flow.fi_init()
    flow.update_moments()
    flow.step(args.steps)
    fi = geo.owned(flow.get_fi(), 1, 3).copy()
    r, u, F = flow.get_arrays()
    r = geo.owned(r, c.stencil_size_rho, 3).copy()
    u = geo.owned(u, 1, 3).copy()
    F = geo.owned(F, 1, 3).copy()
    rhot, prs, velt = flow.update_diagnostics()
    dn0 = flow.delta_norm()
    flow.step(1)
    dn1 = flow.delta_norm()
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(fi=fi, rho=r, u=u, F=F, rhot=rhot, dn=(dn0, dn1)))
    result = {"case": args.case, "world": world, "steps": args.steps}
    if rank == 0:
        G = {k: slab.assemble([g[k] for g in gathered]) for k in ("fi", "rho", "u", "F", "rhot")}
        o = cases.run_oracle_bc(cfg, walls, rho, bcs, args.steps)
        fluid = walls == 0
        result["err_fi"] = float(gpu_util.rel_err(G["fi"], o.fi()))
        result["err_rho"] = float(gpu_util.rel_err(G["rho"][fluid], o.rho()[fluid]))
        result["err_u"] = float(gpu_util.rel_err(G["u"][fluid], o.u()[fluid]))
        den = np.abs(o.forces()).max()
        result["err_F"] = float(np.abs(G["F"] - o.forces()).max() / (den if den > 0 else 1.0))
        result["err_rhot"] = float(gpu_util.rel_err(G["rhot"], o.diagnostics()[0]))
        result["solid_zero"] = bool(np.all(G["fi"][~fluid] == 0.0))
        m0 = gpu_util.mass(np.asarray(rho).reshape(G["rho"].shape), fluid)
        m1 = gpu_util.mass(G["rho"], fluid)
        result["mass_rel"] = 0.0 if bcs else float(np.max(np.abs(m1 - m0) / np.abs(m0)))  # open faces exchange mass
        # the delta norm is a max over ranks in the reference (MPI_Allreduce, lbm_distribution_function.F90:828)
        result["delta_norm_first"] = [g["dn"][0] for g in gathered]
        # fluid-only evaluation: the device stores no solid nodes, where the reference's
        # VecPointwiseDivide computes 0/0
        f0 = o.fi()
        o.step(1)
        f1 = o.fi()
        nz = f1 != 0.0
        result["delta_norm_oracle"] = float(np.abs((f0[nz] - f1[nz]) / f1[nz]).max())
        result["delta_norm_ranks"] = [g["dn"][1] for g in gathered]
        if args.single:
            one = gpu_util.make_flow_bc(cfg, walls, rho, bcs, device=local_rank)
            one.step(args.steps)
            fi1 = geo.owned(one.get_fi(), 1, 3)
            result["bit_identical_to_single_rank"] = bool(np.array_equal(fi1, G["fi"]))
            result["max_abs_vs_single_rank"] = float(np.abs(fi1 - G["fi"]).max())
            one.close()
        Path(args.out).write_text(json.dumps(result))
    dist.barrier()
    flow.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
