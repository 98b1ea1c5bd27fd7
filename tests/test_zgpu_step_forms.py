"""The forms of the order-4 step on the device give the same bits:
  * stage (default for 1, 2 and 4 components): k_step_stage -- populations, adjacency rows and mask of a warp's next item
    fetched by two tensor copies (TMA) into the warp's double buffer in shared memory while the current item is collided
  * clc (TXG_STAGE_CLC=1/0): k_step_stage_clc -- the staged form whose blocks take over the next block of the grid through
    cluster launch control instead of ending, so that every item of a warp but its first is prefetched
  * adjc (TXG_STAGE_ADJC=1/0): k_step_stage fed with one compressed adjacency record per item (bases + byte offsets, 27 instead of
    76 bytes per node) instead of the adjacency and mask rows
  * table (TXG_STAGE=0; the default for 3 and 5 components): k_step_fused (per-node adjacency table, every operand by
    demand loads)
  * band / push (opt-in TXG_BAND=1): k_step_band (bit rows instead of the table, density windows in shared memory)
  * band / pull (opt-in TXG_BAND=1 TXG_PULL=1): k_step_band<PULL> + k_moments_pull -- the population buffer holds collided
    populations between steps and the reference's fi is gathered at the start of the next step, or by k_pull_stream when
    fi itself is asked for (exports, restart, delta norm) in the middle of a run.
Same arithmetic in the same order on the same values, so np.array_equal, not a tolerance; the oracle comparison of the
default form is tests/test_gpu_parity.py."""
import numpy as np
import pytest

import cases
import gpu_util
from taxila_lbm_b200 import config as tc
from taxila_lbm_b200 import geometry as geo

pytestmark = pytest.mark.gpu

FORMS = {"table": dict(TXG_STAGE="0"), "stage": dict(TXG_STAGE_CLC="0", TXG_STAGE_ADJC="0"), "adjc": dict(TXG_STAGE_CLC="0", TXG_STAGE_ADJC="1"),
         "clc": dict(TXG_STAGE_CLC="1"), "push": dict(TXG_BAND="1"), "pull": dict(TXG_BAND="1", TXG_PULL="1")}
KERNEL = {"table": "k_step_fused", "stage": "k_step_stage", "adjc": "k_step_stage", "clc": "k_step_stage_clc", "push": "k_step_band",
          "pull": "k_step_band_pull"}


def run(cfg, walls, rho, form, monkeypatch, chunks, peek=False):
    for k in ("TXG_BAND", "TXG_PULL", "TXG_STAGE", "TXG_STAGE_CLC", "TXG_STAGE_ADJC", "TXG_STAGE_PG", "TXG_STAGE_WARPS", "TXG_BAND_LB", "TXG_LAG", "TXG_RHOTILE", "TXG_SPLIT"):
        monkeypatch.delenv(k, raising=False)
    for k, v in FORMS[form].items():
        monkeypatch.setenv(k, v)
    flow = gpu_util.make_flow(cfg, walls, rho)
    mid = []
    for n in chunks:
        flow.step(n)
        if peek:  # an export in the middle of the run: fi is materialised, the next step starts from streamed populations
            mid.append(gpu_util.fields(flow))
    flow.synchronize()
    out = gpu_util.fields(flow)
    kt = flow.kernel_times()
    flow.close()
    return out, mid, kt


def three_components(N=24):
    cfg = tc.default_config(3, 3, N, N, N)
    for d in range(3):
        cfg.periodic[d] = 1
    for m in range(3):
        for k in range(3):
            if k != m:
                cfg.gf[m][k] = 0.05 + 0.01 * (m + k)
    tc.finalize_flags(cfg)
    walls = geo.porous_spheres(N, N, N, seed=11, rmin=3.0, rmax=6.0, solid_fraction=0.4, nminerals=1)
    rho = 0.2 + 0.6 * np.random.default_rng(5).random((N, N, N, 3))
    rho[walls != 0] = 0.0
    return cfg, walls, rho


CASES = {
    "porous": lambda: cases.porous_3d(32, rmin=4.0, rmax=8.0),
    "porous_wide_rows": lambda: cases.porous_3d(96, 40, 12, rmin=2.0, rmax=4.5),  # three bit-row words per row
    "bubble": lambda: cases.bubble_3d(32),
    "closed": lambda: cases.porous_3d(24, rmin=3.0, rmax=6.0, periodic=(0, 0, 0)),
    "mixed_periodic": lambda: cases.porous_3d(24, rmin=3.0, rmax=6.0, periodic=(1, 0, 1)),
    "x_walls": lambda: cases.porous_3d(24, rmin=3.0, rmax=6.0, periodic=(0, 1, 1)),
    "2d": lambda: cases.bubble_2d(64),
    "2d_mrt": lambda: cases.bubble_2d(48, mrt=True),
    "s3": three_components,
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_forms_bit_identical(monkeypatch, case):
    cfg, walls, rho = CASES[case]()
    steps = 20
    ref, _, k0 = run(cfg, walls, rho, "table", monkeypatch, (steps,))
    assert k0[KERNEL["table"]][1] == steps, k0
    for form in ("stage", "adjc", "clc", "push", "pull"):
        out, _, kt = run(cfg, walls, rho, form, monkeypatch, (3, 1, steps - 4))
        if form in ("stage", "adjc", "clc") and cfg.ncomponents == 3:
            assert kt["k_step_fused"][1] == steps, kt  # (10 positions per item: the staged form is not offered, the table kernel runs)
        else:
            assert kt[KERNEL[form]][1] == steps, (form, kt)
        for a, b, name in zip(ref, out, ("fi", "rho", "u", "forces")):
            assert np.array_equal(a, b), (case, form, name, float(np.abs(a - b).max()))


def test_pull_form_with_exports_in_the_middle(monkeypatch):
    """fi, rho, u, forces read after every chunk: k_pull_stream materialises fi, the following step reads it in place."""
    cfg, walls, rho = cases.porous_3d(32, rmin=4.0, rmax=8.0)
    chunks = (1, 2, 5, 4)
    _, mid0, _ = run(cfg, walls, rho, "table", monkeypatch, chunks, peek=True)
    _, mid1, kt = run(cfg, walls, rho, "pull", monkeypatch, chunks, peek=True)
    assert kt["k_pull_stream"][1] == len(chunks), kt
    for a4, b4 in zip(mid0, mid1):
        for a, b in zip(a4, b4):
            assert np.array_equal(a, b)


def test_pull_form_with_eos(monkeypatch):
    cfg, walls, rho = cases.porous_3d(32, rmin=4.0, rmax=8.0)
    cfg.use_nonideal_eos = 1
    for m in range(2):
        cfg.eos_type[m] = tc.EOS_SC
        cfg.eos_rho0[m] = 0.8 + 0.3 * m
    ref, _, _ = run(cfg, walls, rho, "table", monkeypatch, (25,))
    out, _, kt = run(cfg, walls, rho, "pull", monkeypatch, (25,))
    assert kt["k_step_band_pull"][1] == 25 and kt["k_moments_pull"][1] >= 24, kt
    for a, b in zip(ref, out):
        assert np.array_equal(a, b)


def test_pull_form_delta_norm_and_restart(monkeypatch):
    """DistributionCalcDeltaNorm and a restart from fi (txg_get_fi -> txg_set_fi) in the pull form."""
    cfg, walls, rho = cases.porous_3d(24, rmin=3.0, rmax=6.0)
    vals = {}
    for form in ("table", "pull"):
        for k in ("TXG_BAND", "TXG_PULL", "TXG_STAGE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in FORMS[form].items():
            monkeypatch.setenv(k, v)
        flow = gpu_util.make_flow(cfg, walls, rho)
        flow.step(5)
        n0 = flow.delta_norm()
        flow.step(3)
        n1 = flow.delta_norm()
        fi = flow.get_fi()
        flow.close()
        flow2 = gpu_util.make_flow(cfg, walls, rho)
        flow2.set_fi(fi)
        flow2.update_moments()
        flow2.step(7)
        vals[form] = (n0, n1, fi, gpu_util.fields(flow2))
        flow2.close()
    a, b = vals["table"], vals["pull"]
    assert a[0] == b[0] == 1e99 and a[1] == b[1]
    assert np.array_equal(a[2], b[2])
    for x, y in zip(a[3], b[3]):
        assert np.array_equal(x, y)


def test_staged_form_on_a_box_whose_planes_straddle_items(monkeypatch):
    """k_step_stage on a box whose planes do not start on multiples of 16 positions (items clipped at both ends of a
    launch are replayed lanes), with short and long blocks, and on a box of 10 k items."""
    cfg, walls, rho = cases.porous_3d(40, 24, 20, rmin=3.0, rmax=6.0)
    ref, _, _ = run(cfg, walls, rho, "table", monkeypatch, (15,))
    out, _, kt = run(cfg, walls, rho, "stage", monkeypatch, (15,))
    assert kt["k_step_stage"][1] == 15, kt
    monkeypatch.setenv("TXG_STAGE_ROUNDS", "5")  # longer blocks: every warp refills its stages several times
    out5, _, _ = run(cfg, walls, rho, "stage", monkeypatch, (15,))
    monkeypatch.delenv("TXG_STAGE_ROUNDS")
    for a, b in zip(ref, out5):
        assert np.array_equal(a, b)
    for a, b in zip(ref, out):
        assert np.array_equal(a, b)
    outc, _, kt = run(cfg, walls, rho, "clc", monkeypatch, (15,))
    assert kt["k_step_stage_clc"][1] == 15, kt
    for a, b in zip(ref, outc):
        assert np.array_equal(a, b)
    # 10 k items: more blocks than are resident at once, so that the blocks of the clc form do take over later ones
    cfg, walls, rho = cases.porous_3d(96, 96, 40, rmin=4.0, rmax=9.0)
    ref, _, _ = run(cfg, walls, rho, "table", monkeypatch, (6,))
    for form in ("stage", "adjc", "clc"):
        out, _, kt = run(cfg, walls, rho, form, monkeypatch, (2, 4))
        for a, b in zip(ref, out):
            assert np.array_equal(a, b), form


def test_compressed_adjacency_escape_items(monkeypatch):
    """TXG_STAGE_ADJC=1 on boxes whose rows are longer than 256 fluid nodes and periodic in x: the items at the two ends of a row
    hold neighbour positions a whole row apart, their byte offsets overflow and the lanes fall back to the full table; and on a box
    with clipped items at both ends of the owned range."""
    for cfg, walls, rho in (cases.bubble_3d(288, 4), cases.porous_3d(40, 24, 20, rmin=3.0, rmax=6.0)):
        ref, _, _ = run(cfg, walls, rho, "table", monkeypatch, (8,))
        out, _, kt = run(cfg, walls, rho, "adjc", monkeypatch, (3, 5))
        assert kt["k_step_stage"][1] == 8, kt
        for a, b in zip(ref, out):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("form", ["stage", "adjc", "clc"])
def test_staged_forms_with_gather_prefetch(monkeypatch, form):
    """TXG_STAGE_PG=1: a warp prefetches the densities and the wall record of its next item into L1 while it collides the current one
    (it waits for the next item's adjacency in the middle of the current item): same bits, blocks of 1, 2 and 5 rounds."""
    cfg, walls, rho = cases.porous_3d(96, 96, 40, rmin=4.0, rmax=9.0)
    ref, _, _ = run(cfg, walls, rho, "table", monkeypatch, (6,))
    for rounds in (1, 2, 5):
        FORMS["_pg"] = dict(FORMS[form], TXG_STAGE_PG="1", TXG_STAGE_ROUNDS=str(rounds))
        try:
            out, _, kt = run(cfg, walls, rho, "_pg", monkeypatch, (2, 4))
        finally:
            del FORMS["_pg"]
            monkeypatch.delenv("TXG_STAGE_ROUNDS", raising=False)
        assert kt[KERNEL[form]][1] == 6, kt
        for a, b in zip(ref, out):
            assert np.array_equal(a, b), (form, rounds)


@pytest.mark.parametrize("warps", [6, 12])
def test_staged_block_shapes(monkeypatch, warps):
    """The staged K2 with 6 warps x 2 blocks and 12 warps x 1 block per SM (TXG_STAGE_WARPS; room for L1), static blocks and
    blocks that take over the next one (TXG_STAGE_CLC=1), short and long blocks: the same bits as the table kernel."""
    cfg, walls, rho = cases.porous_3d(96, 96, 40, rmin=4.0, rmax=9.0)
    ref, _, _ = run(cfg, walls, rho, "table", monkeypatch, (6,))
    for form, rounds in (("stage", 2), ("stage", 5), ("clc", 1), ("clc", 4)):
        FORMS["_shape"] = dict(FORMS[form], TXG_STAGE_WARPS=str(warps), TXG_STAGE_ROUNDS=str(rounds))
        try:
            out, _, kt = run(cfg, walls, rho, "_shape", monkeypatch, (2, 4))
        finally:
            del FORMS["_shape"]
            monkeypatch.delenv("TXG_STAGE_ROUNDS", raising=False)
        assert kt[KERNEL[form]][1] == 6, kt
        for a, b in zip(ref, out):
            assert np.array_equal(a, b), (form, rounds)


@pytest.mark.parametrize("case", ["porous_iso8", "closed_iso8", "hots_2d_iso10", "bubble_2d_iso8", "s3_iso8"])
def test_wide_stencil_forms_bit_identical(monkeypatch, case):
    """Orders 8 and 10: k_forces_tile (forces out of a dense shared-memory tile of psi) + k_collide (the default) ==
    k_step_tile (the same, then collide + push in one kernel; opt-in TXG_WIDE_FUSED=1) == the map-walking k_forces +
    k_collide (TXG_FORCES_TILE=0), bit for bit."""
    if case == "porous_iso8":
        cfg, walls, rho = cases.porous_3d(40, 24, 20, order=8, rmin=3.0, rmax=6.0)
    elif case == "closed_iso8":
        cfg, walls, rho = cases.porous_3d(24, order=8, rmin=3.0, rmax=6.0, periodic=(0, 1, 0))
    elif case == "hots_2d_iso10":
        cfg, walls, rho = cases.bubble_2d_hots(96)
    elif case == "bubble_2d_iso8":
        cfg, walls, rho = cases.bubble_2d(70, mrt=True, order=8)
    else:
        cfg, walls, rho = three_components(20)
        cfg.isotropy_order = 8
        cfg.stencil_size_rho = 2
        tc.finalize_flags(cfg)
    steps = 12
    outs = {}
    for name, env, kernel in (("tile_fused", dict(TXG_WIDE_FUSED="1"), "k_step_tile"), ("tile_split", {}, "k_forces_tile"),
                              ("map_split", dict(TXG_FORCES_TILE="0"), "k_forces")):
        for k in ("TXG_SPLIT", "TXG_FORCES_TILE", "TXG_WIDE_FUSED", "TXG_BAND", "TXG_PULL", "TXG_STAGE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        flow = gpu_util.make_flow(cfg, walls, rho)
        flow.step(steps)
        outs[name] = gpu_util.fields(flow)
        kt = flow.kernel_times()
        flow.close()
        assert kt[kernel][1] >= steps, (name, kt)
    for name in ("tile_split", "tile_fused"):
        for a, b, what in zip(outs["map_split"], outs[name], ("fi", "rho", "u", "forces")):
            assert np.array_equal(a, b), (case, name, what, float(np.abs(a - b).max()))
