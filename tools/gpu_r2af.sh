#!/bin/bash
# usage: gpurun --timeout 400 -- tools/gpu_r2af.sh   (how much of K2 is the MRT transform: SRT ablation, with and without stores)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache TXG_BENCH_NOCHECK=1
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2af_$n.json 2> gpurun_out/r2af_$n.err || tail -3 gpurun_out/r2af_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2af_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"] and n.startswith("k_")}, d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["clocks"]["reasons"])
PY
}
P=$PWD/taxila-lbm_b200
run mrt
run srt TXG_BENCH_SRT=1
run srt_no_stores TXG_BENCH_SRT=1 TAXILA_GPU_LIB=$P/libtaxila_gpu_abl3.so
run srt_aligned_stores TXG_BENCH_SRT=1 TAXILA_GPU_LIB=$P/libtaxila_gpu_abl2.so
