#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
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
run w8 libtaxila_gpu.so
run w8_o6 libtaxila_gpu.so TXG_OPTS=6
run w8_o4 libtaxila_gpu.so TXG_OPTS=4
run w8_c100 libtaxila_gpu.so TXG_CARVEOUT=100
run w4 libtaxila_gpu_w4.so
run plain libtaxila_gpu.so TXG_STREAM=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 3 -c 1 -o gpurun_out/collide_stream_w8 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_collide_stream_w8.log 2>&1; echo "ncu rc=$?"
