#!/usr/bin/env python3
"""MLUPS of the CPU oracle (oracle/, the reference's algorithm restated in C with OpenMP) on the configurations BASELINE.json names,
at sizes the oracle finishes in seconds.  TEST INFRASTRUCTURE (runs the oracle); run where no GPU is needed:
    python tools/cpu_oracle_configs.py > profiles/r2_cpu_oracle_configs.txt"""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402
import oracle  # noqa: E402

CASES = [
    ("C1 tests/bubble_2D as shipped (D2Q9, SRT, 128^2)", lambda: cases.bubble_2d(), 200),
    ("C2 tests/bubble_2D_hots (D2Q9, MRT, order 10, 128^2)", lambda: cases.bubble_2d_hots(), 100),
    ("C3 tests/bubble_3D as shipped (D3Q19, SRT, 128^3)", lambda: cases.bubble_3d(128), 6),
    ("C4 recipe, 128^3 crop (D3Q19, MRT, 3 minerals, body force, 45 % fluid)", lambda: cases.porous_3d(128), 6),
    ("C4 recipe, order-8 stencil, 128^3 crop", lambda: cases.porous_3d(128, order=8), 4),
]

print("# CPU oracle on %d cores of the BUILD container (not the GPU box: its 16 cores do 9.3-9.9 MLUPS on the 256^3 crop of C4, profiles/r2final_bench_reference.json)" % os.cpu_count())
print("# MLUPS = all nodes of the box (solid ones included, like the GPU metric) x steps / wall time of the steps; set-up excluded")
for name, build, steps in CASES:
    cfg, walls, rho = build()
    nodes = cfg.NX * cfg.NY * cfg.NZ
    for threads in (1, os.cpu_count()):
        o = oracle.Oracle(cfg, threads=threads)
        o.set_walls(walls)
        o.set_rho(rho)
        o.fi_init()
        o.update_moments()
        o.step(1)
        t0 = time.time()
        o.step(steps)
        dt = time.time() - t0
        print("%-75s threads %2d  %8.2f ms/step  %7.2f MLUPS" % (name, threads, dt / steps * 1e3, nodes * steps / dt / 1e6))
        del o
