#!/bin/bash
# usage: gpurun --timeout 600 -- tools/gpu_r2ad.sh   (where do the L2 sectors of K2 go: base vs 12-warp blocks vs clc, selected counters)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__m_xbar2l1tex_read_sectors.sum,lts__t_bytes.sum.per_second,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__d_sectors_fill_sysmem.sum,lts__t_sectors_srcnode_gpc.sum,lts__t_sectors_srcunit_l1.sum
cap() { # name kernel env...
  n=$1; k=$2; shift; shift
  env "$@" timeout 250 ncu --metrics $M --clock-control none -k regex:$k -s 4 -c 1 --csv --log-file gpurun_out/r2ad_$n.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2> gpurun_out/r2ad_$n.err
  tail -2 gpurun_out/r2ad_$n.err
}
cap base 'k_step_stage$'
cap w12_r4 'k_step_stage$' TXG_STAGE_WARPS=12 TXG_STAGE_ROUNDS=4
cap w12_r8 'k_step_stage$' TXG_STAGE_WARPS=12 TXG_STAGE_ROUNDS=8
cap moments 'k_moments'
ls -la gpurun_out/r2ad_*
