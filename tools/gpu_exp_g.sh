#!/bin/bash
source tools/gpu_try.sh
run slabs8 libtaxila_gpu.so
run slabs1 libtaxila_gpu.so TXG_SLABS=1
run slabs4 libtaxila_gpu.so TXG_SLABS=4
run slabs16 libtaxila_gpu.so TXG_SLABS=16
run slabs32 libtaxila_gpu.so TXG_SLABS=32
