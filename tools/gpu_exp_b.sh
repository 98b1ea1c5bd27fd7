#!/bin/bash
source tools/gpu_try.sh
export TXG_BENCH_NOCHECK=1
run base libtaxila_gpu.so
run abl1_nostores libtaxila_gpu_abl1.so
run abl2_nomath libtaxila_gpu_abl2.so
run abl3_nogather libtaxila_gpu_abl3.so
run abl5_copy libtaxila_gpu_abl5.so
