#!/bin/bash
mkdir -p gpurun_out
run() { # name lib pf
  TAXILA_GPU_LIB=$PWD/taxila-lbm_b200/$2 TXG_PF=$3 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/x_$1.json 2> gpurun_out/x_$1.err || tail -3 gpurun_out/x_$1.err
  python - $1 <<'PY'
import json,sys
d=json.load(open("gpurun_out/x_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]})
PY
}
run base_pf0 libtaxila_gpu.so 0
run base_pf592 libtaxila_gpu.so 592
run mb5_pf0 libtaxila_gpu_mb5.so 0
run mb6_pf0 libtaxila_gpu_mb6.so 0
run mb6_pf592 libtaxila_gpu_mb6.so 888
TXG_PF=592 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 3 -c 1 -o gpurun_out/collide_pf592_512 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_collide_pf.log 2>&1; echo "ncu rc=$?"
