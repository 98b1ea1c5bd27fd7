"""Oracle parity at the sizes the metric is quoted on: the size-dependent code of the device
path -- 32-bit push indices, bit-row adjacency with many words per row, density windows that span several rows, the
periodic-x wrap look-ups, the L2 prefetch past a block -- is compared value by value with the CPU oracle, not only
through mass conservation.

  * C4 recipe (D3Q19 two-component MRT, 3 minerals, body force, spheres r = 10..22) at the 128^3 crop x 1000 steps
    (the parity bar of the headline config),
  * C3 tests/bubble_3D as shipped: 128^3, SRT, 100 steps (tests/bubble_3D/initialize_state.F90:104-122),
  * C3 scaled to 256^3 x 50 steps (BASELINE.json configs[2]).

The oracle runs on all host cores (OpenMP); about 5 minutes of CPU work in total.  Each case appends one JSON line to
$TXG_PARITY_LOG when that is set (tools/gpu_r2_evidence.sh keeps it under profiles/).
"""
import json
import os
import time

import numpy as np
import pytest

import cases
import gpu_util

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

TOL = 1e-10  # max |gpu - oracle| / max |oracle| per field (north star: 1e-10 after 1000 steps)


def _compare(name, cfg, walls, rho, steps):
    threads = os.cpu_count() or 1
    t0 = time.time()
    o = cases.run_oracle(cfg, walls, rho, steps, threads=threads)
    t_oracle = time.time() - t0
    t0 = time.time()
    flow = gpu_util.make_flow(cfg, walls, rho)
    flow.step(steps)
    fi, r, u, F = gpu_util.fields(flow)
    t_gpu = time.time() - t0
    fluid = np.asarray(walls).reshape(r.shape[:3]) == 0
    errs = {
        "fi": gpu_util.rel_err(fi, o.fi()),
        "rho": gpu_util.rel_err(r[fluid], o.rho()[fluid]),
        "u": gpu_util.rel_err(u[fluid], o.u()[fluid]),
        "forces": gpu_util.rel_err(F[fluid], o.forces()[fluid]),
    }
    m0 = gpu_util.mass(np.asarray(rho).reshape(r.shape), fluid)
    m1 = gpu_util.mass(r, fluid)
    drift = float(np.max(np.abs(m1 - m0) / np.abs(m0)))
    kernels = {k: int(v[1]) for k, v in flow.kernel_times().items()}
    flow.close()
    rec = {"case": name, "box": [cfg.NX, cfg.NY, cfg.NZ], "steps": steps, "fluid_nodes": int(fluid.sum()),
           "max_rel_err": {k: float(v) for k, v in errs.items()}, "mass_drift_rel": drift, "tol": TOL,
           "oracle_s": round(t_oracle, 1), "oracle_threads": threads, "gpu_s_incl_transfers": round(t_gpu, 1), "kernels": kernels}
    print(json.dumps(rec))
    if os.environ.get("TXG_PARITY_LOG"):
        with open(os.environ["TXG_PARITY_LOG"], "a") as fh:
            fh.write(json.dumps(rec) + "\n")
    for k, v in errs.items():
        assert v <= TOL, (k, v, errs)
    assert drift <= 1e-12, drift
    assert np.all(fi[~fluid] == 0.0) and np.all(r[~fluid] == 0.0)


def test_c4_recipe_128_cubed_1000_steps():
    _compare("C4 porous MRT 3 minerals body force, 128^3 crop", *cases.porous_3d(128), steps=1000)


def test_bubble_3d_as_shipped_128_cubed_100_steps():
    _compare("C3 bubble_3D as shipped (SRT)", *cases.bubble_3d(128), steps=100)


def test_bubble_3d_256_cubed_50_steps():
    _compare("C3 bubble_3D scaled to 256^3 (SRT)", *cases.bubble_3d(256, hw=20), steps=50)


# ---------------------------------------------------------------------------------------------------------------------------------
# 256^3 x 1000 steps against committed samples of the oracle's result (tests/golden/make_large_golden.py made them in the build
# container: 1.5 h of oracle time per case does not fit the GPU box's clock).  Same deterministic builders, same seeds.
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _compare_golden(name, stem, cfg, walls, rho, steps):
    path = os.path.join(GOLD, stem + ".npz")
    if not os.path.exists(path):
        pytest.skip("no golden sample %s (tests/golden/make_large_golden.py)" % stem)
    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    assert meta["box"] == [cfg.NX, cfg.NY, cfg.NZ] and meta["steps"] == steps, meta
    t0 = time.time()
    flow = gpu_util.make_flow(cfg, walls, rho)
    flow.step(steps)
    fi, r, u, F = gpu_util.fields(flow)
    t_gpu = time.time() - t0
    n = np.asarray(walls).size
    idx = g["idx"]
    fluid = np.asarray(walls).reshape(-1) == 0
    assert fluid[idx].all()
    errs = {
        "rho": float(np.abs(r.reshape(n, -1)[idx] - g["rho"]).max() / g["rho_max"].max()),
        "u": float(np.abs(u.reshape(n, -1)[idx] - g["u"]).max() / g["u_max"]),
        "fi": float(np.abs(fi.reshape(n, -1)[idx[:g["fi"].shape[0]]] - g["fi"]).max() / g["fi_max"]),
    }
    fl3 = fluid.reshape(r.shape[:3])
    m0 = gpu_util.mass(np.asarray(rho).reshape(r.shape), fl3)
    m1 = gpu_util.mass(r, fl3)
    drift = float(np.max(np.abs(m1 - m0) / np.abs(m0)))
    kernels = {k: int(v[1]) for k, v in flow.kernel_times().items()}
    flow.close()
    rec = {"case": name, "box": [cfg.NX, cfg.NY, cfg.NZ], "steps": steps, "against": "tests/golden/%s.npz" % stem, "sampled_nodes": int(idx.size),
           "max_rel_err": errs, "mass_drift_rel": drift, "tol": TOL, "oracle_s": meta["oracle_s"], "oracle_threads": meta["threads"],
           "gpu_s_incl_transfers": round(t_gpu, 1), "kernels": kernels}
    print(json.dumps(rec))
    if os.environ.get("TXG_PARITY_LOG"):
        with open(os.environ["TXG_PARITY_LOG"], "a") as fh:
            fh.write(json.dumps(rec) + "\n")
    for k, v in errs.items():
        assert v <= TOL, (k, v, errs)
    assert drift <= 1e-12, drift


def test_bubble_3d_256_cubed_1000_steps_vs_golden_sample():
    """BASELINE.json configs[2] at its full size and the north star's 1000 steps."""
    _compare_golden("C3 bubble_3D scaled to 256^3 (SRT), 1000 steps", "c3_bubble3d_256_1000", *cases.bubble_3d(256, hw=20), steps=1000)


def test_c4_recipe_256_cubed_1000_steps_vs_golden_sample():
    """The C4 recipe on the 256^3 sample SURVEY.md 8d names, 1000 steps."""
    _compare_golden("C4 porous MRT 3 minerals body force, 256^3", "c4_porous_256_1000", *cases.porous_3d(256), steps=1000)
