#!/bin/bash
mkdir -p gpurun_out
export TXG_BENCH_NOCHECK=1
run() { # name lib env
  n=$1; lib=$2; shift; shift
  env TAXILA_GPU_LIB=$PWD/taxila-lbm_b200/$lib "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/x_$n.json 2> gpurun_out/x_$n.err || tail -3 gpurun_out/x_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/x_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]})
PY
}
run plain libtaxila_gpu.so TXG_STREAM=0
run abl4_alignedstores libtaxila_gpu_abl4.so TXG_STREAM=0
run abl5_copy libtaxila_gpu_abl5.so TXG_STREAM=0
