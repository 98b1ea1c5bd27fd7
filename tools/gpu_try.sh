#!/bin/bash
# usage: tools/gpu_try.sh  -- parity tests, then bench.py kernel times for a list of (name, lib, env...) variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { # name lib env...
  n=$1; lib=$2; shift; shift
  env TAXILA_GPU_LIB=$PWD/taxila-lbm_b200/$lib "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/x_$n.json 2> gpurun_out/x_$n.err || tail -3 gpurun_out/x_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/x_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
