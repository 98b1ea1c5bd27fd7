#!/usr/bin/env python3
"""bench.py -- MLUPS of the Taxila-LBM flow hot path on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W              # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU arm

Workload at N=1: BASELINE.json configs[3] -- D3Q19 two-component Shan-Chen, MRT, 512^3 random
overlapping-sphere porous medium (three minerals with their own wettability), body force,
bounce-back walls, flushing initial state, periodic box.
N>1 (weak scaling, configs[4]): the same 512^3 block per GPU, stacked along z (the global
geometry is the C4 box tiled N times along the periodic z axis, so every halo is a real NCCL
exchange between different GPUs); --scaling strong splits the one 512^3 box instead.

A "step" is one LBM time step of the whole box.  value = NX*NY*NZ_global*K / t / 1e6 with t the
max over ranks of the CUDA-event time of the K steps (state resident in HBM).  e2e = the same
metric through the reference-facing call sequence with HOST buffers: walls + initial densities
uploaded, FlowFiInit, K steps issued as the six LBMRun2 procedure calls per step, and the
FlowUpdateDiagnostics fields (rhot, prs, velt) copied back -- all inside the timed region.

The CPU arm (--impl reference) times the reference's algorithm restated in C (oracle/, OpenMP over
all host cores -- the reference itself needs a Fortran compiler, PETSc 3.6 and MPI, none of which
exist in this image) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import numpy as np  # noqa: E402

METRIC = "MLUPS"
B_ALG_FLUID = 641.0  # 16*S*Q + 16*S + 1 for D3Q19, S=2
B_ALG_SOLID = 1.0
# algorithmic bytes per fluid node of each hot kernel (each datum once; the adjacency tables of the sparse
# storage are NOT counted: they are overhead of this implementation and show up in `traffic`)
B_K2_FLUID = 625.0  # fused forces+collide kernel: 16*S*Q (f in, f out) + 8*S (rho) + 1 (node class)   [SURVEY 8d]
B_K2B_FLUID = 657.0  # split collide kernel: 16*S*Q + 8*S*D (forces in) + 1
B_K1_FLUID = 320.0  # moments kernel: 8*S*Q + 8*S
B_KF_FLUID = 65.0   # split forces kernel: 8*S (rho) + 8*S*D (F out) + 1


def bind_near_gpu(index):
    """Run this rank (and the page-locked host arrays it touches first) on the CPUs next to its GPU: with 8 ranks on a
    two-socket box the host <-> device copies of the e2e leg otherwise cross the socket link.  Host-side placement only."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return "cpus %d" % len(os.sched_getaffinity(0))
    except Exception as e:  # not fatal: placement is an optimisation
        return "unbound (%s)" % type(e).__name__


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "power_w_max": float(max(pw)) if pw else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


_WALLS = {}


def build_case(size, nranks, rank, scaling, order, nz=None):
    """Per-rank slab of the workload: (cfg, walls_rg, rho_rg, fluid_fraction, global_nodes)."""
    from taxila_lbm_b200 import geometry as geo
    from taxila_lbm_b200 import slab, workloads

    # (sweeps of several bench runs in one GPU call keep the generated geometry: TXG_CASE_CACHE=<dir>)
    cache = os.environ.get("TXG_CASE_CACHE")
    nz = size if nz is None else nz
    cpath = Path(cache) / ("c4_%d_%d_walls.npy" % (size, nz)) if cache else None
    walls = _WALLS.get((size, nz))  # (the strong-scaling sub-measurement re-uses the geometry of the weak one)
    if walls is None and cpath is not None and cpath.exists():
        walls = np.load(cpath)
    have = walls is not None
    # TXG_BENCH_SRT=1: timing ablation only (how much of K2 is the MRT transform); the line then says relaxation SRT
    cfg, walls, rho = workloads.porous_3d(size, size, nz, order=order, walls=walls, mrt=os.environ.get("TXG_BENCH_SRT") != "1")
    _WALLS[(size, nz)] = walls
    if cpath is not None and not have and rank == 0:
        np.save(cpath, walls)
    R = cfg.stencil_size_rho
    if nranks == 1 or scaling == "strong":
        # the one size^3 box, split into z-slabs (DMDA ownership ranges)
        c, walls_rg, rho_rg = slab.local_arrays(cfg, walls, rho, nranks, rank)
        NZg = nz
    else:
        # weak: the global box is the size^3 block tiled nranks times along periodic z and rank r
        # owns tile r, so every halo is a real exchange between different GPUs.  All tiles hold the
        # same geometry, so the ghost planes of a tile are its own opposite planes; the flushing IC
        # puts the invading fluid in the first 10 planes of the GLOBAL box only, i.e. in tile 0.
        NZg = nz * nranks
        c = cfg.copy()
        c.NZ = NZg
        c.zs, c.zl = rank * nz, nz
        c.rank, c.nranks = rank, nranks
        if rank != 0:
            rho = geo.flushing_rho(cfg, walls, (0.03, 0.97), (0.03, 0.97), "z", 10)
        walls_rg = geo.ghosted(walls, R, cfg.periodic, 3, wall_ghost=True)
        rho_rg = geo.ghosted(rho, R, cfg.periodic, 3)  # ghost planes of rho are refreshed by the halo exchange
    fluid_local = float((geo.owned(walls_rg, R, 3) == 0).mean())
    return c, walls_rg, rho_rg, fluid_local, size * size * NZg


def run_cpu_sample(order, sample_size, steps, threads, warm=1):
    """The oracle (reference structure, OpenMP) on a sample_size^3 crop of the same recipe: `warm` untimed steps
    (page faults, caches), then exactly `steps` timed ones.  Returns (MLUPS, seconds per step)."""
    import oracle  # tests/oracle.py: the ctypes wrapper of oracle/ (CPU legs only)
    from taxila_lbm_b200 import workloads

    cfg, walls, rho = workloads.porous_3d(sample_size, order=order)
    o = oracle.Oracle(cfg, threads=threads)
    o.set_walls(walls)
    o.set_rho(rho)
    o.fi_init()
    o.update_moments()
    o.step(max(1, warm))
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    o.close()
    return sample_size ** 3 * steps / dt / 1e6, dt / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512, help="block edge (512 = BASELINE config)")
    ap.add_argument("--nz", type=int, default=None, help="planes of the box (default: --size); a thin box is the per-GPU work of strong scaling")
    ap.add_argument("--order", type=int, default=4, help="isotropy order of the Shan-Chen stencil")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-sample", type=int, default=256, help="edge of the crop the CPU legs time (SURVEY.md 8d: 256^3)")
    ap.add_argument("--cpu-steps", type=int, default=10, help="timed oracle steps of the cpu_baseline leg of the CUDA arm")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="N > 1, weak scaling: skip the strong-scaling sub-measurement")
    args = ap.parse_args()

    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner under
    # NCCL_DEBUG=VERSION, extension build chatter) are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)
    workload = "D3Q19 two-component Shan-Chen MRT, %d^3 random-sphere porous medium per GPU, 3 minerals, body force, " \
               "bounce-back, iso-%d" % (args.size, args.order)
    if os.environ.get("TXG_BENCH_SRT") == "1":
        workload = "ABLATION (SRT instead of MRT): " + workload
    config = {"workload": workload, "lattice": "D3Q19", "components": 2, "relaxation": "SRT" if os.environ.get("TXG_BENCH_SRT") == "1" else "MRT",
              "box_per_gpu": [args.size, args.size, args.nz or args.size], "isotropy_order": args.order, "geometry": "porous_spheres(seed=20260)",
              "l2_policy": "inputs larger than L2 (40.8 GB of populations per 512^3 block)",
              "cpu_sample_box": [args.cpu_sample] * 3}

    # ------------------------------------------------------------------ CPU arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        # exactly `warmup` untimed and `steps` timed LBM steps of the crop: a step of this arm is one time step of the
        # bounded sample (config.cpu_sample_box), not of the 512^3 box -- MLUPS is size-normalised
        v, sps = run_cpu_sample(args.order, args.cpu_sample, args.steps, threads, warm=warmup)
        sample = "%d^3 crop of the same porous recipe (config.cpu_sample_box): %d untimed + %d timed LBM steps of the crop, " \
                 "%d OpenMP threads" % (args.cpu_sample, warmup, args.steps, threads)
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warmup, "ms_per_step": sps * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "timed": {"box": [args.cpu_sample] * 3, "steps": args.steps, "warmup": warmup, "ms_per_step_of_that_box": sps * 1e3},
            "cpu_baseline": {"value": v, "unit": "MLUPS", "cores": threads, "kind": "port", "sample": sample,
                             "note": "C restatement of the reference's CPU algorithm in the reference's structure "
                                     "(oracle/), OpenMP; the Fortran+PETSc reference cannot be built in this image"},
            "e2e": {"value": v, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit(line)
        return 0

    # ------------------------------------------------------------------ CUDA arm
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_near_gpu(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import taxila_lbm_b200 as tx
    from taxila_lbm_b200 import geometry as geo

    cfg, walls_rg, rho_rg, fluid_frac, global_nodes = build_case(args.size, world, rank, args.scaling, args.order, args.nz)
    nccl_id = None
    if world > 1:
        ids = [tx.Flow.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]
    flow = tx.Flow(cfg, device=local_rank, nccl_id=nccl_id)
    flow.walls_set_values(walls_rg)
    flow.initialize_state(rho_rg)
    flow.fi_init()
    flow.update_moments()
    flow.step(warmup)
    flow.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    # the timed region: exactly K steps as the application issues them (one txg_step(K) call; the library replays pairs of
    # steps as CUDA graphs), between barriers, timed by CUDA events on the handle's stream
    barrier()
    t_host = time.perf_counter()
    flow.step(args.steps)
    host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / max(args.steps, 1)  # (the call returns when the work is queued)
    flow.synchronize()
    barrier()
    ms, launches = flow.last_step_ms()
    # the same K steps once more with a CUDA-event pair around every kernel launch (eager launches: events cannot sit
    # inside a replayed graph): the per-kernel durations of `roofline` and `kernels`
    flow.reset_kernel_times()
    flow.enable_kernel_timing(True)
    barrier()
    flow.step(args.steps)
    flow.synchronize()
    barrier()
    ms_timed_pass, _ = flow.last_step_ms()
    ktimes = flow.kernel_times()
    flow.enable_kernel_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = global_nodes * args.steps / (ms_max * 1e-3) / 1e6

    # sanity: the state is finite and the mass of each component over ALL ranks is what the initial
    # state held (no work skipped, no population lost in a halo)
    R = cfg.stencil_size_rho
    fluid = geo.owned(walls_rg, R, 3) == 0
    rho_now = geo.owned(flow.get_arrays(u=False, forces=False)[0], R, 3)
    nocheck = os.environ.get("TXG_BENCH_NOCHECK") == "1"  # ablation builds only
    assert np.isfinite(rho_now).all() or nocheck
    m = torch.tensor([[float(np.sum(a[..., k][fluid], dtype=np.longdouble)) for k in range(cfg.ncomponents)]
                      for a in (geo.owned(rho_rg, R, 3), rho_now)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(m, op=dist.ReduceOp.SUM)
    mass_drift = float(((m[1] - m[0]).abs() / m[0].abs()).max().item())
    assert mass_drift <= 1e-10 or nocheck, mass_drift

    # ------------------------------------------------------------------ e2e through the host-buffer API
    e2e = None
    if not args.no_e2e:
        # the application's own host arrays, page-locked (allocated once, outside the timed region, like
        # the reference's PETSc Vecs): inputs are copied host -> device and results device -> host inside it
        def pinned(shape):
            return torch.empty(tuple(shape), dtype=torch.float64, pin_memory=True).numpy()

        walls_h, rho_h = pinned(walls_rg.shape), pinned(rho_rg.shape)
        walls_h[...] = walls_rg
        rho_h[...] = rho_rg
        diag_h = tuple(pinned(sh) for sh in flow.shape_diagnostics())
        walls_rg, rho_rg = walls_h, rho_h
        barrier()
        t0 = time.perf_counter()
        marks = [("start", t0)]

        sync_marks = os.environ.get("TXG_BENCH_SYNC_MARKS") == "1"  # diagnosis only: wait for the device at every mark

        def mark(name):  # (host clock; the calls before a mark are synchronous except the steps, which the last mark covers)
            if sync_marks:
                flow.synchronize()
            marks.append((name, time.perf_counter()))

        flow.walls_set_values(walls_rg)
        mark("walls_set_values")
        flow.initialize_state(rho_rg)
        mark("initialize_state")
        flow.fi_init()
        if sync_marks:
            mark("fi_init")
        flow.update_moments()
        mark("fi_init+update_moments (queued)")
        for _ in range(args.steps):
            flow.collision(); flow.communicate_fi(); flow.stream(); flow.bounceback(); flow.apply_bcs(); flow.update_flux()
        mark("steps (queued)")
        out = flow.update_diagnostics(out=diag_h)
        flow.synchronize()
        mark("update_diagnostics incl. the wait for the steps")
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = (walls_rg.nbytes + rho_rg.nbytes) * world
        d2h = sum(a.nbytes for a in out) * world
        e2e = {"value": global_nodes * args.steps / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "what": "walls+rho upload, FlowFiInit, FlowUpdateMoments, %d steps as the six LBMRun2 procedure calls, "
                       "FlowUpdateDiagnostics fields copied back into page-locked host arrays; host wall clock, max over ranks" % args.steps,
               "host_placement": numa,
               "breakdown_ms": {marks[i][0]: round((marks[i][1] - marks[i - 1][1]) * 1e3, 2) for i in range(1, len(marks))}}

    # ------------------------------------------------------------------ strong scaling beside the weak line (N > 1)
    # the ONE size^3 box split into N z-slabs, same timed protocol (warm-up, K steps between barriers, CUDA events, max
    # over ranks): the latency-bound case of SURVEY.md 8e (512 / N planes per GPU)
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_strong:
        flow.close()
        cfg_s, walls_s, rho_s, _, nodes_s = build_case(args.size, world, rank, "strong", args.order)
        ids = [tx.Flow.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        fs_ = tx.Flow(cfg_s, device=local_rank, nccl_id=ids[0])
        fs_.walls_set_values(walls_s)
        fs_.initialize_state(rho_s)
        fs_.fi_init()
        fs_.update_moments()
        fs_.step(warmup)
        fs_.synchronize()
        barrier()
        fs_.step(args.steps)
        fs_.synchronize()
        barrier()
        ms_s, launches_s = fs_.last_step_ms()
        ts = torch.tensor([ms_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        r_s = geo.owned(fs_.get_arrays(u=False, forces=False)[0], cfg_s.stencil_size_rho, 3)
        assert np.isfinite(r_s).all()
        fs_.close()
        ms_s = float(ts.item())
        strong = {"value": nodes_s * args.steps / (ms_s * 1e-3) / 1e6, "unit": "MLUPS", "ms_per_step": ms_s / args.steps,
                  "box": [args.size] * 3, "planes_per_gpu": cfg_s.zl, "steps": args.steps, "warmup": warmup,
                  "launches_per_step_per_rank": launches_s / max(args.steps, 1),
                  "what": "the one %d^3 box of the N=1 line split into %d z-slabs; efficiency = value / (N x the N=1 value)" % (args.size, world)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ roofline of the dominant kernel
    peak, peak_src = measured_peak()
    nodes_local = args.size ** 2 * cfg.zl
    # dominant kernel: the fused forces+collide kernel (order-4 stencil) or the split collide kernel
    dom, dom_bytes = ("k_step_fused", B_K2_FLUID) if ktimes.get("k_step_fused", (0.0, 0))[1] else ("k_collide", B_K2B_FLUID)
    if dom == "k_collide":  # split path (wide stencils, face BCs): whichever of its two kernels takes longer
        for alt in ("k_forces", "k_forces_tile"):
            if ktimes.get(alt, (0.0, 0))[0] > ktimes.get(dom, (0.0, 0))[0]:
                dom, dom_bytes = alt, B_KF_FLUID
    if ktimes.get("k_step_fused_tile", (0.0, 0))[1]:  # TXG_RHOTILE=1 (opt-in): same bytes as k_step_fused
        dom, dom_bytes = "k_step_fused_tile", B_K2_FLUID
    for alt in ("k_step_stage", "k_step_stage_clc", "k_step_band", "k_step_band_pull"):  # other forms of K2: same algorithmic bytes as k_step_fused
        if ktimes.get(alt, (0.0, 0))[1]:
            dom, dom_bytes = alt, B_K2_FLUID
    if ktimes.get("k_step_fused_lag", (0.0, 0))[1]:  # TXG_LAG=1 (opt-in one-pass step): the whole step's bytes in one launch
        dom, dom_bytes = "k_step_fused_lag", B_ALG_FLUID
    kc_ms, kc_n = ktimes.get(dom, (0.0, 0))
    roofline = None
    if kc_n:
        # per-launch algorithmic bytes: launches may cover sub-ranges of the slab (boundary/interior
        # split); bytes of all launches of one step add up to the slab
        per_step_bytes = nodes_local * (fluid_frac * dom_bytes + (1 - fluid_frac) * B_ALG_SOLID)
        achieved = per_step_bytes * args.steps / (kc_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tj = ROOT / "profiles" / "traffic.json"
        if tj.exists():
            try:
                t = json.loads(tj.read_text()).get(dom)
                if t and t.get("size") == args.size and t.get("order") == args.order and world == 1:
                    traffic, traffic_src = t["dram_bytes_per_launch"], t["source"]
            except Exception:
                pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "bytes_per_fluid_node": dom_bytes, "bytes_per_launch": per_step_bytes * args.steps / kc_n,
                    "avg_launch_ms": kc_ms / kc_n, "share_of_step": kc_ms / (ms_timed_pass if ms_timed_pass else 1.0),
                    "timed_in": "a second pass of the same %d steps with an event pair around every launch" % args.steps}
    step_bytes = global_nodes * (fluid_frac * B_ALG_FLUID + (1 - fluid_frac) * B_ALG_SOLID)
    step_roofline = {"bytes_per_lup_fluid": B_ALG_FLUID, "fluid_fraction": fluid_frac,
                     "achieved_gbs_per_gpu": step_bytes * args.steps / (ms_max * 1e-3) / 1e9 / world,
                     "frac_of_hbm_peak": step_bytes * args.steps / (ms_max * 1e-3) / 1e9 / world / peak,
                     "mflups": value * fluid_frac}
    kernels = {k: {"ms": v[0], "launches": v[1]} for k, v in ktimes.items()}
    # per-kernel achieved algorithmic GB/s (the launches of one step add up to the slab)
    for name, b in (("k_moments", B_K1_FLUID), ("k_forces", B_KF_FLUID), ("k_forces_tile", B_KF_FLUID), ("k_collide", B_K2B_FLUID), ("k_step_fused", B_K2_FLUID),
                    ("k_step_fused_tile", B_K2_FLUID), ("k_step_band", B_K2_FLUID), ("k_step_band_pull", B_K2_FLUID),
                    ("k_step_stage", B_K2_FLUID), ("k_step_stage_clc", B_K2_FLUID), ("k_moments_pull", B_K1_FLUID), ("k_step_fused_lag", B_ALG_FLUID)):
        if name == "k_moments" and "k_step_fused_lag" in kernels and kernels["k_step_fused_lag"]["launches"]:
            continue  # one-pass step: k_moments only sums the two boundary planes
        if name in kernels and kernels[name]["ms"] > 0:
            by = nodes_local * fluid_frac * b * args.steps
            kernels[name]["algorithmic_gbs"] = by / (kernels[name]["ms"] * 1e-3) / 1e9
            kernels[name]["frac_of_peak"] = kernels[name]["algorithmic_gbs"] / peak

    cpu = None
    if not args.no_cpu and world == 1:  # (the CPU leg beside the N = 1 line only; the CPU arm proper is --impl reference)
        os.sched_setaffinity(0, all_cpus)  # the CPU leg gets every host core, as in --impl reference
        threads = os.cpu_count() or 1
        v, sps = run_cpu_sample(args.order, args.cpu_sample, args.cpu_steps, threads, warm=2)
        cpu = {"value": v, "unit": "MLUPS", "cores": threads, "kind": "port",
               "sample": "%d^3 crop of the same porous recipe, 2 untimed + %d timed steps, oracle/ (reference structure, OpenMP, "
                         "%d threads), %.1f s" % (args.cpu_sample, args.cpu_steps, threads, sps * args.cpu_steps)}

    line = {
        "metric": METRIC, "value": value, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(launches) * world, "host_enqueue_ms_per_step": host_enqueue_ms,
        "ms_per_step_with_kernel_events": ms_timed_pass / args.steps, "mass_drift_rel": mass_drift, "roofline": roofline, "step_roofline": step_roofline, "kernels": kernels,
        "cpu_baseline": cpu,
    }
    if strong is not None:
        line["strong"] = strong
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
