#!/bin/bash
# usage: gpurun --timeout 900 -- tools/gpu_r2y.sh   (A/B of the staged and the table K2 on one box, alternating, with the clocks of each run)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
run() { # name args env...
  n=$1; a=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu $a > gpurun_out/r2y_$n.json 2> gpurun_out/r2y_$n.err || tail -3 gpurun_out/r2y_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2y_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"] and n.startswith("k_")}, d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["clocks"]["reasons"])
PY
}
nvidia-smi --query-gpu=name,power.limit,clocks.sm,clocks.max.sm --format=csv
run stage1 ""
run table1 "" TXG_STAGE=0
run stage2 ""
run table2 "" TXG_STAGE=0
run stage3 ""
run table3 "" TXG_STAGE=0
TXG_RUN_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zzz_experimental_lag.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -3
