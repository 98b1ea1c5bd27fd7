"""The reference's shipped regression test (tests/bubble_2D) replayed from compiled C through the C ABI
(shim/replay_bubble_2d.c): the call sequence of an ISO_C_BINDING shim, no Python in the loop."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def build(tmp_path):
    exe = tmp_path / "replay_bubble_2d"
    lib = ROOT / "taxila-lbm_b200"
    r = subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "shim" / "replay_bubble_2d.c"),
                        "-L", str(lib), "-ltaxila_gpu", "-Wl,-rpath," + str(lib), "-lm", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_harness_compiles_and_fails_loudly_without_a_gpu(tmp_path):
    import torch

    exe = build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([str(exe), str(ROOT / "tests" / "golden" / "bubble_2D_fi001.dat")], capture_output=True, text=True)
    assert r.returncode == 2
    assert "no CPU path" in r.stderr


@pytest.mark.gpu
def test_bubble_2d_regression_through_the_c_abi(tmp_path):
    exe = build(tmp_path)
    r = subprocess.run([str(exe), str(ROOT / "tests" / "golden" / "bubble_2D_fi001.dat")], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS (round-off)" in r.stdout
