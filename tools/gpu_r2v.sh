#!/bin/bash
# usage: gpurun --timeout 1200 -- tools/gpu_r2v.sh   (z-marching tile forces: parity of the wide-stencil forms, 512^3 order-8 timings)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
( time timeout 900 python -m pytest tests/test_zgpu_step_forms.py tests/test_gpu_parity.py tests/test_zgpu_face_bcs.py -q -m "gpu and not slow" --tb=short -p no:cacheprovider -k "wide or iso or hots or reflecting or five or diagnostics" ) > gpurun_out/r2v_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2v_pytest_gpu.log
run() { # name args env...
  n=$1; a=$2; shift; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu $a > gpurun_out/r2v_$n.json 2> gpurun_out/r2v_$n.err || tail -3 gpurun_out/r2v_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2v_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run iso8_m16 "--order 8"
run iso8_m32 "--order 8" TXG_TILE_MARCH=32
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_forces_tile -s 4 -c 1 -o gpurun_out/r2v_forces_tile python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --order 8 > /dev/null 2> gpurun_out/r2v_forces_ncu.err
ls -la gpurun_out/r2v*.ncu-rep
