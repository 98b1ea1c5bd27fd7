#!/bin/bash
# usage: gpurun --timeout 500 -- tools/gpu_r2ak.sh   (fused diagnostics export + FlowFiInit without fp64 divisions: tests, e2e with device-synchronised marks, plain e2e)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_zgpu_restart.py tests/test_c_harness.py -q -m "gpu and not slow" --tb=short -p no:cacheprovider -k "diagnostics or golden or output_files or bubble_2d_vs_oracle or porous_mrt or eos or restart or six_procedure or c_abi" 2>&1 | tail -6
TXG_BENCH_SYNC_MARKS=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2ak_sync.json 2> gpurun_out/r2ak_sync.err || tail -3 gpurun_out/r2ak_sync.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2ak_bench20.json 2> gpurun_out/r2ak_bench20.err || tail -3 gpurun_out/r2ak_bench20.err
python - <<'PY'
import json
for n in ("sync", "bench20"):
    d=json.load(open("gpurun_out/r2ak_%s.json" % n))
    print(n, "value %.0f ms/step %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["e2e"]["breakdown_ms"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
