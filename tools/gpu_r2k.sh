#!/bin/bash
# usage: gpurun --timeout 1500 -- tools/gpu_r2k.sh   (full fast GPU suite; staged K2 (static short blocks, tensor copies); fused wide-stencil step)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
nvidia-smi -L > gpurun_out/r2k_pytest_gpu.log
( time timeout 1200 python -m pytest tests -q -m "gpu and not slow" --tb=short -p no:cacheprovider ) >> gpurun_out/r2k_pytest_gpu.log 2>&1
tail -8 gpurun_out/r2k_pytest_gpu.log
run() { # name args env...
  n=$1; a=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu $a > gpurun_out/r2k_$n.json 2> gpurun_out/r2k_$n.err || tail -3 gpurun_out/r2k_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2k_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run stage "" TXG_STAGE=1
run table ""
run stage_r3 "" TXG_STAGE=1 TXG_STAGE_ROUNDS=3
run stage_b "" TXG_STAGE=1
run table_b ""
run iso8_fused "--order 8"
run iso8_split "--order 8" TXG_SPLIT=1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_step_tile -s 4 -c 1 -o gpurun_out/r2k_step_tile python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --order 8 > /dev/null 2> gpurun_out/r2k_tile_ncu.err
ls -la gpurun_out/r2k*.ncu-rep
