"""Build libtaxila_gpu.so (nvcc, sm_100a only) in-tree."""
import os
import subprocess
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"


def build(jobs=None, verbose=False):
    jobs = jobs or os.cpu_count() or 4
    env = dict(os.environ)
    # the image's CC/CXX point at a wrapper without OpenMP specs; nvcc wants the system g++
    env.pop("CC", None)
    env.pop("CXX", None)
    r = subprocess.run(["make", "-C", str(CSRC), "-j%d" % jobs], capture_output=not verbose, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc build failed:\n%s\n%s" % (r.stdout or "", r.stderr or ""))
    lib = CSRC.parent / "libtaxila_gpu.so"
    assert lib.exists()
    return lib
