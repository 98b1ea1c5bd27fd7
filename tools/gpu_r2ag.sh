#!/bin/bash
# usage: gpurun --timeout 600 -- tools/gpu_r2ag.sh   (compressed adjacency records of the staged K2: bit-identity tests, A/B on one box)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
timeout 400 python -m pytest tests/test_zgpu_step_forms.py -q -m gpu --tb=short -p no:cacheprovider -k "forms_bit_identical or straddle or escape" 2>&1 | tail -8
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2ag_$n.json 2> gpurun_out/r2ag_$n.err || tail -3 gpurun_out/r2ag_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2ag_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"] and n.startswith("k_")}, d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["clocks"]["reasons"], "drift", d.get("mass_drift_rel"))
PY
}
run adjc1 TXG_STAGE_ADJC=1
run base1 TXG_STAGE_ADJC=0
run adjc2 TXG_STAGE_ADJC=1
run base2 TXG_STAGE_ADJC=0
run adjc_r3 TXG_STAGE_ADJC=1 TXG_STAGE_ROUNDS=3
