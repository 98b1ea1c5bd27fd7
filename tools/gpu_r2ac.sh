#!/bin/bash
# usage: gpurun --timeout 700 -- tools/gpu_r2ac.sh   (staged K2 in wider blocks -- 6 warps x 2, 12 warps x 1 per SM, L1 left for the density rows -- static and clc)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
timeout 300 python -m pytest tests/test_zgpu_step_forms.py -q -m gpu --tb=short -p no:cacheprovider -k "block_shapes" 2>&1 | tail -8
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2ac_$n.json 2> gpurun_out/r2ac_$n.err || tail -3 gpurun_out/r2ac_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2ac_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"] and n.startswith("k_")}, d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["clocks"]["reasons"])
PY
}
run base
run w12_r2 TXG_STAGE_WARPS=12 TXG_STAGE_ROUNDS=2
run w12_r4 TXG_STAGE_WARPS=12 TXG_STAGE_ROUNDS=4
run w12_r8 TXG_STAGE_WARPS=12 TXG_STAGE_ROUNDS=8
run w12_clc_r2 TXG_STAGE_WARPS=12 TXG_STAGE_ROUNDS=2 TXG_STAGE_CLC=1
run w12_clc_r4 TXG_STAGE_WARPS=12 TXG_STAGE_ROUNDS=4 TXG_STAGE_CLC=1
run w6_r4 TXG_STAGE_WARPS=6 TXG_STAGE_ROUNDS=4
run w6_r8 TXG_STAGE_WARPS=6 TXG_STAGE_ROUNDS=8
run w6_clc_r2 TXG_STAGE_WARPS=6 TXG_STAGE_ROUNDS=2 TXG_STAGE_CLC=1
run w6_clc_r4 TXG_STAGE_WARPS=6 TXG_STAGE_ROUNDS=4 TXG_STAGE_CLC=1
run w12_r4_c100 TXG_STAGE_WARPS=12 TXG_STAGE_ROUNDS=4 TXG_STAGE_CARVE=100
