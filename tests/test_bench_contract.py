"""bench.py without a GPU: the CPU arm (--impl reference) prints ONE JSON line with the contract's keys and says what it
actually timed; the CUDA arm refuses to run without a device (the hot path has no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3", "--cpu-sample", "64"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MLUPS" and d["unit"] == "MLUPS" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    # what was timed is in the line: the crop, its steps, the threads
    assert d["timed"]["box"] == [64, 64, 64] and d["timed"]["steps"] == 2
    assert d["config"]["cpu_sample_box"] == [64, 64, 64] and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "64^3" in cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_cuda_arm_refuses_without_a_device():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--size", "16"], capture_output=True, text=True,
                       timeout=600, cwd=str(ROOT))
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
