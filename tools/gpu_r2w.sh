#!/bin/bash
# usage: gpurun --timeout 900 -- tools/gpu_r2w.sh   (plain L2 prefetch distance of the staged K2: TXG_STAGE_PF < 0)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
run() { # name args env...
  n=$1; a=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu $a > gpurun_out/r2w_$n.json 2> gpurun_out/r2w_$n.err || tail -3 gpurun_out/r2w_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2w_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"] and n.startswith("k_")}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run pf0 ""
run pfm74 "" TXG_STAGE_PF=-74
run pfm148 "" TXG_STAGE_PF=-148
run pfm296 "" TXG_STAGE_PF=-296
run pfm592 "" TXG_STAGE_PF=-592
