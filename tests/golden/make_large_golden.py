"""Golden samples of the CPU oracle at the sizes BASELINE.json names, 1000 steps (north star: fields within 1e-10 after 1000 steps):

    python tests/golden/make_large_golden.py bubble256   # C3 tests/bubble_3D scaled to 256^3 (SRT), 1000 steps
    python tests/golden/make_large_golden.py porous256   # C4 recipe (MRT, 3 minerals, body force) at 256^3, 1000 steps

An oracle run of this size takes 0.5 - 1.5 h on 8 cores -- too long for the GPU box's clock -- so it is made HERE, once, and what is
committed is a sample of its result: rho and u at 16384 seeded fluid nodes, fi at the first 2048 of them, the mass per component.
tests/test_zgpu_large_parity.py rebuilds the same inputs from the same deterministic builders, runs the device path and compares at the
sampled nodes.  TEST INFRASTRUCTURE (runs the oracle)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

import cases  # noqa: E402

NS, NFI, SEED = 16384, 2048, 424242


def sample_nodes(walls):
    fluid = np.flatnonzero(np.asarray(walls).reshape(-1) == 0)
    rng = np.random.default_rng(SEED)
    return np.sort(rng.choice(fluid, size=min(NS, fluid.size), replace=False))


CASES = {
    "bubble256": (lambda: cases.bubble_3d(256, hw=20), 1000, "c3_bubble3d_256_1000"),
    "porous256": (lambda: cases.porous_3d(256), 1000, "c4_porous_256_1000"),
}


def main(which, threads=8):
    build, steps, stem = CASES[which]
    cfg, walls, rho = build()
    t0 = time.time()
    o = cases.run_oracle(cfg, walls, rho, steps, threads=threads)
    secs = time.time() - t0
    idx = sample_nodes(walls)
    r, u, fi = o.rho(), o.u(), o.fi()
    fluid = np.asarray(walls).reshape(-1) == 0
    r2 = r.reshape(fluid.size, -1)
    out = dict(idx=idx, rho=r2[idx], u=u.reshape(fluid.size, -1)[idx], fi=fi.reshape(fluid.size, -1)[idx[:NFI]], mass=r2[fluid].sum(axis=0),
               rho_max=np.abs(r2).max(axis=0), u_max=np.abs(u).max(), fi_max=np.abs(fi).max(),
               meta=json.dumps(dict(case=which, box=[cfg.NX, cfg.NY, cfg.NZ], steps=steps, seed=SEED, oracle_s=round(secs, 1), threads=threads,
                                    fi_layout=list(fi.shape))))
    np.savez_compressed(HERE / (stem + ".npz"), **out)
    print(which, "done in %.0f s" % secs, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k != "meta"})


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 8)
