#!/bin/bash
source tools/gpu_try.sh
run base libtaxila_gpu.so
run b3 libtaxila_gpu_b3.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 3 -c 1 -o gpurun_out/r1e_collide -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r1e_ncu_collide.log 2>&1; echo "ncu rc=$?"
