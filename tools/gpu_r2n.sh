#!/bin/bash
# usage: gpurun --timeout 900 -- tools/gpu_r2n.sh   (L2 tensor prefetch distance of the staged K2)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
( time timeout 600 python -m pytest tests/test_zgpu_step_forms.py -q -m "gpu and not slow" --tb=short -p no:cacheprovider -k "forms_bit_identical or staged" ) > gpurun_out/r2n_pytest_gpu.log 2>&1
tail -4 gpurun_out/r2n_pytest_gpu.log
run() { # name args env...
  n=$1; a=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu $a > gpurun_out/r2n_$n.json 2> gpurun_out/r2n_$n.err || tail -3 gpurun_out/r2n_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2n_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run pf0 ""
run pf74 "" TXG_STAGE_PF=74
run pf148 "" TXG_STAGE_PF=148
run pf296 "" TXG_STAGE_PF=296
run pf592 "" TXG_STAGE_PF=592
run pf148_r3 "" TXG_STAGE_PF=148 TXG_STAGE_ROUNDS=3
