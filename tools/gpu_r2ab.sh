#!/bin/bash
# usage: gpurun --timeout 600 -- tools/gpu_r2ab.sh   (k_step_stage_clc: bit-identity tests, then A/B against the static staged K2 on one box)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
timeout 300 python -m pytest tests/test_zgpu_step_forms.py -q -m gpu --tb=short -p no:cacheprovider -k "forms_bit_identical or straddle" 2>&1 | tail -8
run() { # name args env...
  n=$1; a=$2; shift; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu $a > gpurun_out/r2ab_$n.json 2> gpurun_out/r2ab_$n.err || tail -3 gpurun_out/r2ab_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2ab_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"] and n.startswith("k_")}, d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["clocks"]["reasons"], "drift", d.get("mass_drift_rel"))
PY
}
run clc1 "" TXG_STAGE_CLC=1
run stage1 "" TXG_STAGE_CLC=0
run clc2 "" TXG_STAGE_CLC=1
run stage2 "" TXG_STAGE_CLC=0
run clc_r1 "" TXG_STAGE_CLC=1 TXG_STAGE_ROUNDS=1
run clc_r4 "" TXG_STAGE_CLC=1 TXG_STAGE_ROUNDS=4
