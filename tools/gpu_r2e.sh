#!/bin/bash
# usage (one gpurun call, ~20 GPU-minutes): gpurun --timeout 1800 -- tools/gpu_r2e.sh
# The staged form of K2 (k_step_stage, default) and the pull form (opt-in) on a B200: -m gpu suite, 512^3 timings, ncu.
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
nvidia-smi -L > gpurun_out/r2e_pytest_gpu.log
( time timeout 1200 python -m pytest tests/test_zgpu_step_forms.py tests/test_gpu_parity.py tests/test_zgpu_face_bcs.py -q -m "gpu and not slow" --tb=short -p no:cacheprovider ) >> gpurun_out/r2e_pytest_gpu.log 2>&1
tail -12 gpurun_out/r2e_pytest_gpu.log
run() { # name env...
  n=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2e_$n.json 2> gpurun_out/r2e_$n.err || tail -3 gpurun_out/r2e_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2e_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run stage
run table TXG_STAGE=0
run stage_c4 TXG_STAGE_CHUNKS=4
run stage_c64 TXG_STAGE_CHUNKS=64
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_step_stage -s 4 -c 1 -o gpurun_out/r2e_stage python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2> gpurun_out/r2e_stage_ncu.err
tail -3 gpurun_out/r2e_stage_ncu.err
ls -la gpurun_out/r2e_stage.ncu-rep
