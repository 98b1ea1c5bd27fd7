#!/bin/bash
# usage: gpurun --timeout 600 -- tools/gpu_r2ae.sh   (store ablations of the staged K2: timing only, TXG_BENCH_NOCHECK)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache TXG_BENCH_NOCHECK=1
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2ae_$n.json 2> gpurun_out/r2ae_$n.err || tail -3 gpurun_out/r2ae_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2ae_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"] and n.startswith("k_")}, d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["clocks"]["reasons"])
PY
}
P=$PWD/taxila-lbm_b200
run base
run abl1_no_bounce_stores TAXILA_GPU_LIB=$P/libtaxila_gpu_abl1.so
run abl2_aligned_stores TAXILA_GPU_LIB=$P/libtaxila_gpu_abl2.so
run abl3_no_stores TAXILA_GPU_LIB=$P/libtaxila_gpu_abl3.so
run base2
