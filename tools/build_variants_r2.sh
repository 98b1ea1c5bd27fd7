#!/bin/bash
# usage (here, before a gpurun call): tools/build_variants_r2.sh -- cache-policy builds of the population traffic (see build_variants.sh)
set -e
cd "$(dirname "$0")/../taxila-lbm_b200/csrc"
unset CC CXX
build() { make -j"$(nproc)" OBJDIR=build_$1 TARGET=../libtaxila_gpu_$1.so EXTRA="$2" > /dev/null; echo "built libtaxila_gpu_$1.so ($2)"; }
build ldna "-DTXG_LDF_MODE=1"
build stcs "-DTXG_STF_MODE=1"
build stcg "-DTXG_STF_MODE=2"
build ldna_stcs "-DTXG_LDF_MODE=1 -DTXG_STF_MODE=1"
