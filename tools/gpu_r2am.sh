#!/bin/bash
# usage: gpurun --timeout 200 -- tools/gpu_r2am.sh   (full ncu capture of k_export_diag_fused in the e2e leg of bench.py; a first run used
# -k regex:'k_export_diag_fused|k_fi_init_fused' -c 2 -o gpurun_out/r2am_e2e_kernels and caught the two FlowFiInit launches)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_CASE_CACHE=/tmp/txg_cache
timeout 180 ncu --set full --clock-control none --import-source on -k regex:k_export_diag_fused -c 1 -o gpurun_out/r2am_export_diag -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2am_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2am_ncu.log; ls -la gpurun_out/r2am_export_diag.ncu-rep
