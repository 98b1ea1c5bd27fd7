#!/bin/bash
source tools/gpu_try.sh
run base libtaxila_gpu.so
