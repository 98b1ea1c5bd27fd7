#!/bin/bash
# usage (here, before a gpurun call): tools/build_variants.sh
# Builds kernel-tuning variants of libtaxila_gpu.so next to the default one (git-ignored, they travel with gpurun):
#   libtaxila_gpu_ldna.so   population loads without L1 allocation      (-DTXG_LDF_MODE=1)
#   libtaxila_gpu_stcs.so   streaming population stores                 (-DTXG_STF_MODE=1)
#   libtaxila_gpu_stcg.so   st.global.cg population stores              (-DTXG_STF_MODE=2)
#   libtaxila_gpu_ldna_stcs.so  both
#   libtaxila_gpu_laghints.so   L2 residency hints of the one-pass step (-DTXG_LAG_HINTS=1)
#   libtaxila_gpu_cap128.so / cap256.so   window length of k_step_fused_tile (-DTXG_TILE_CAP=128 / 256)
set -e
cd "$(dirname "$0")/../taxila-lbm_b200/csrc"
unset CC CXX
build() { # name flags
  make -j"$(nproc)" OBJDIR=build_$1 TARGET=../libtaxila_gpu_$1.so EXTRA="$2" > /dev/null
  echo "built libtaxila_gpu_$1.so ($2)"
}
build ldna "-DTXG_LDF_MODE=1"
build stcs "-DTXG_STF_MODE=1"
build stcg "-DTXG_STF_MODE=2"
build ldna_stcs "-DTXG_LDF_MODE=1 -DTXG_STF_MODE=1"
build cap128 "-DTXG_TILE_CAP=128"   # window of k_step_fused_tile (TXG_RHOTILE=1; default 192)
build cap256 "-DTXG_TILE_CAP=256"
build laghints "-DTXG_LAG_HINTS=1"   # one-pass step: input evict-first, pushed populations evict-last until summed
