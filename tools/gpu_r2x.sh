#!/bin/bash
# usage: gpurun --gpus 2 --timeout 1200 -- tools/gpu_r2x.sh   (face BCs on the fused step: BC / EOS / restart tests, 2-rank BC case, a 256^3 drainage timing fused vs split)
mkdir -p gpurun_out
export TXG_MG_LOG=$PWD/gpurun_out/r2x_parity_mg_results.jsonl
rm -f $TXG_MG_LOG
( time timeout 900 python -m pytest tests/test_zgpu_face_bcs.py tests/test_zgpu_eos.py tests/test_zgpu_restart.py tests/test_zz_multi_gpu_bcs.py tests/test_c_harness.py -m gpu -v --tb=short -p no:cacheprovider ) > gpurun_out/r2x_pytest_bc.log 2>&1
grep -n "PASSED\|FAILED\|ERROR\|passed\|failed" gpurun_out/r2x_pytest_bc.log | tail -30
python - <<'PY' > gpurun_out/r2x_bc_timing.txt 2>&1
import os, sys, time
sys.path.insert(0, "tests")
import numpy as np
import cases, gpu_util
from taxila_lbm_b200 import config as tc
for env in ({}, {"TXG_SPLIT": "1"}):
    os.environ.pop("TXG_SPLIT", None)
    os.environ.update(env)
    c, walls, rho, bcs = cases.drainage_3d(N=256, NZ=256, x_bc=None)
    flow = gpu_util.make_flow_bc(c, walls, rho, bcs)
    flow.step(5); flow.synchronize()
    flow.reset_kernel_times(); flow.enable_kernel_timing(True)
    flow.step(10); flow.synchronize()
    ms, _ = flow.last_step_ms()
    print(env or "fused", "ms/step %.3f" % (ms / 10), {k: round(v[0] / max(v[1], 1), 3) for k, v in flow.kernel_times().items() if v[1]})
    flow.close()
PY
cat gpurun_out/r2x_bc_timing.txt | tail -4
