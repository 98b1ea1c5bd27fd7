"""EXPERIMENTAL, opt-in (TXG_RUN_EXPERIMENTAL=1): the one-pass step (TXG_LAG=1, csrc/lag_schedule.h + k_step_fused_lag)
and the shared-memory density tiles (TXG_RHOTILE=1, k_step_fused_tile) against the oracle and, bit for bit, against the
default two-kernel step.  The kernel was written in a session
without GPU minutes and has not run on a GPU yet, so these tests stay out of the default `-m gpu` run (they sort
last and skip unless asked for); tools/gpu_lag_try.sh runs them first and then times the variants."""
import os

import numpy as np
import pytest

import cases
import gpu_util

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TXG_RUN_EXPERIMENTAL") != "1", reason="experimental kernel: TXG_RUN_EXPERIMENTAL=1")]


def run(cfg, walls, rho, steps, env, monkeypatch, chunks=(None,)):
    for k in ("TXG_LAG", "TXG_LAG_ROWS", "TXG_LAG_PLANES", "TXG_LAG_MPOS", "TXG_RHOTILE"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, str(v))
    flow = gpu_util.make_flow(cfg, walls, rho)
    done = 0
    for c in chunks:  # several txg_step calls: the density must stay current across them
        n = steps - done if c is None else c
        flow.step(n)
        done += n
    flow.synchronize()
    out = gpu_util.fields(flow) + (flow.kernel_times(),)
    flow.close()
    return out


@pytest.mark.parametrize("rows,planes,mpos", [(128, 1, 512), (2, 0, 64), (8, 3, 128), (5, 1, 4096)])
def test_lag_step_equals_default_step_and_oracle(monkeypatch, rows, planes, mpos):
    cfg, walls, rho = cases.porous_3d(32, rmin=4.0, rmax=8.0)
    steps = 20
    fi0, r0, u0, F0, k0 = run(cfg, walls, rho, steps, {}, monkeypatch)
    assert "k_step_fused_lag" not in k0 or k0["k_step_fused_lag"][1] == 0
    fi1, r1, u1, F1, k1 = run(cfg, walls, rho, steps, dict(TXG_LAG=1, TXG_LAG_ROWS=rows, TXG_LAG_PLANES=planes, TXG_LAG_MPOS=mpos),
                              monkeypatch, chunks=(3, 1, None))
    assert k1["k_step_fused_lag"][1] == steps, k1
    assert np.array_equal(fi0, fi1) and np.array_equal(r0, r1) and np.array_equal(u0, u1)
    o = cases.run_oracle(cfg, walls, rho, steps)
    fluid = walls == 0
    assert gpu_util.rel_err(fi1, o.fi()) <= 1e-10
    assert gpu_util.rel_err(r1[fluid], o.rho()[fluid]) <= 1e-10


def test_lag_step_without_solids_and_in_a_closed_box(monkeypatch):
    """no solid node at all (identity position map), SRT; and a box with non-periodic faces (999 ghost walls)"""
    for case in (cases.bubble_3d(32), cases.porous_3d(24, mrt=False, rmin=3.0, rmax=6.0, periodic=(0, 0, 0)),
                 cases.porous_3d(24, rmin=3.0, rmax=6.0, periodic=(1, 0, 1))):
        cfg, walls, rho = case
        fi0, r0, u0, F0, k0 = run(cfg, walls, rho, 15, {}, monkeypatch)
        fi1, r1, u1, F1, k1 = run(cfg, walls, rho, 15, dict(TXG_LAG=1, TXG_LAG_ROWS=8), monkeypatch)
        assert k1["k_step_fused_lag"][1] == 15, k1
        assert np.array_equal(fi0, fi1) and np.array_equal(r0, r1)


def test_lag_step_1000_steps(monkeypatch):
    cfg, walls, rho = cases.porous_3d(48, order=4, rmin=4.0, rmax=9.0)
    fi1, r1, u1, F1, k1 = run(cfg, walls, rho, 1000, dict(TXG_LAG=1, TXG_LAG_ROWS=16), monkeypatch)
    o = cases.run_oracle(cfg, walls, rho, 1000)
    fluid = walls == 0
    assert gpu_util.rel_err(r1[fluid], o.rho()[fluid]) <= 1e-10
    assert gpu_util.rel_err(u1[fluid], o.u()[fluid]) <= 1e-10


def test_lag_step_two_ranks(monkeypatch, tmp_path):
    """z-slabs: the one-pass step on two ranks against the oracle and bit for bit against one rank (both with TXG_LAG=1)."""
    import test_multi_gpu as mg

    if mg.ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("TXG_LAG", "1")
    monkeypatch.setenv("TXG_LAG_ROWS", "8")
    for case, steps in (("porous_periodic", 30), ("porous_closed_box", 30)):
        mg.check(mg.run_case(case, 2, steps, tmp_path, 29631))


@pytest.mark.parametrize("case", ["porous", "bubble", "closed", "2d", "s3"])
def test_rho_tile_kernel_equals_default_step(monkeypatch, case):
    """TXG_RHOTILE=1 (k_step_fused_tile: neighbour densities staged in shared memory by bulk copies, global fallback
    outside the window): bit for bit the default fused step."""
    from taxila_lbm_b200 import config as tc
    from taxila_lbm_b200 import geometry as geo

    if case == "porous":
        cfg, walls, rho = cases.porous_3d(32, rmin=4.0, rmax=8.0)
    elif case == "bubble":
        cfg, walls, rho = cases.bubble_3d(32)
    elif case == "closed":
        cfg, walls, rho = cases.porous_3d(24, rmin=3.0, rmax=6.0, periodic=(0, 0, 0))
    elif case == "2d":
        cfg, walls, rho = cases.bubble_2d(64)
    else:
        cfg = tc.default_config(3, 3, 24, 24, 24)
        for d in range(3):
            cfg.periodic[d] = 1
        for m in range(3):
            for k in range(3):
                if k != m:
                    cfg.gf[m][k] = 0.05 + 0.01 * (m + k)
        tc.finalize_flags(cfg)
        walls = geo.porous_spheres(24, 24, 24, seed=11, rmin=3.0, rmax=6.0, solid_fraction=0.4, nminerals=1)
        rho = 0.2 + 0.6 * np.random.default_rng(5).random((24, 24, 24, 3))
        rho[walls != 0] = 0.0
    steps = 20
    fi0, r0, u0, F0, k0 = run(cfg, walls, rho, steps, {}, monkeypatch)
    fi1, r1, u1, F1, k1 = run(cfg, walls, rho, steps, dict(TXG_RHOTILE=1), monkeypatch)
    assert k1["k_step_fused_tile"][1] == steps, k1
    assert np.array_equal(fi0, fi1) and np.array_equal(r0, r1) and np.array_equal(u0, u1)


def test_lag_step_with_density_tiles(monkeypatch):
    """TXG_LAG=1 TXG_RHOTILE=1: the C blocks of the one-pass step also stage the neighbour densities in shared memory."""
    for case in (cases.porous_3d(32, rmin=4.0, rmax=8.0), cases.bubble_3d(32),
                 cases.porous_3d(24, rmin=3.0, rmax=6.0, periodic=(0, 0, 0))):
        cfg, walls, rho = case
        fi0, r0, u0, F0, k0 = run(cfg, walls, rho, 20, {}, monkeypatch)
        fi1, r1, u1, F1, k1 = run(cfg, walls, rho, 20, dict(TXG_LAG=1, TXG_RHOTILE=1, TXG_LAG_ROWS=8), monkeypatch, chunks=(2, None))
        assert k1["k_step_fused_lag"][1] == 20, k1
        assert np.array_equal(fi0, fi1) and np.array_equal(r0, r1) and np.array_equal(u0, u1)
