#!/bin/bash
source tools/gpu_try.sh
run base libtaxila_gpu.so
run b3 libtaxila_gpu_b3.so
run pf592 libtaxila_gpu.so TXG_PF=592
run pf1184 libtaxila_gpu.so TXG_PF=1184
run pf2368 libtaxila_gpu.so TXG_PF=2368
run b3_pf888 libtaxila_gpu_b3.so TXG_PF=888
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 3 -c 1 -o gpurun_out/r1b_collide -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r1b_ncu_collide.log 2>&1; echo "ncu rc=$?"
