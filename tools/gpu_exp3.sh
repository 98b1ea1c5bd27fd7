#!/bin/bash
mkdir -p gpurun_out
export TXG_BENCH_NOCHECK=1
run() { # name lib pf
  TAXILA_GPU_LIB=$PWD/taxila-lbm_b200/$2 TXG_PF=$3 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/x_$1.json 2> gpurun_out/x_$1.err || tail -3 gpurun_out/x_$1.err
  python - $1 <<'PY'
import json,sys
d=json.load(open("gpurun_out/x_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]})
PY
}
run base libtaxila_gpu.so 0
run abl1_nostores libtaxila_gpu_abl1.so 0
run abl2_nocollide libtaxila_gpu_abl2.so 0
run abl3_nopsi libtaxila_gpu_abl3.so 0
