#!/bin/bash
# usage (here, before a gpurun call): tools/build_variants_r2b.sh -- store ablations of the staged K2 (timing only: the results are wrong)
set -e
cd "$(dirname "$0")/../taxila-lbm_b200/csrc"
unset CC CXX
build() { make -j"$(nproc)" OBJDIR=build_$1 TARGET=../libtaxila_gpu_$1.so EXTRA="$2" > /dev/null; echo "built libtaxila_gpu_$1.so ($2)"; }
build abl1 "-DTXG_ABL_STORE=1"
build abl2 "-DTXG_ABL_STORE=2"
build abl3 "-DTXG_ABL_STORE=3"
