#!/bin/bash
# usage: gpurun --timeout 700 -- tools/gpu_r2ah.sh   (TXG_STAGE_PG=1: gathers of a warp's next item prefetched into L1; tests, sweep, counters)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
timeout 400 python -m pytest tests/test_zgpu_step_forms.py -q -m gpu --tb=short -p no:cacheprovider -k "gather_prefetch" 2>&1 | tail -8
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2ah_$n.json 2> gpurun_out/r2ah_$n.err || tail -3 gpurun_out/r2ah_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2ah_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"] and n.startswith("k_")}, d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["clocks"]["reasons"])
PY
}
run base
run pg_r2 TXG_STAGE_PG=1
run pg_r4 TXG_STAGE_PG=1 TXG_STAGE_ROUNDS=4
run pg_adjc_r3 TXG_STAGE_PG=1 TXG_STAGE_ADJC=1 TXG_STAGE_ROUNDS=3
run pg_clc_r1 TXG_STAGE_PG=1 TXG_STAGE_CLC=1 TXG_STAGE_ROUNDS=1
run pg_w12_clc_r1 TXG_STAGE_PG=1 TXG_STAGE_CLC=1 TXG_STAGE_ROUNDS=1 TXG_STAGE_WARPS=12
run pg_w12_clc_r2 TXG_STAGE_PG=1 TXG_STAGE_CLC=1 TXG_STAGE_ROUNDS=2 TXG_STAGE_WARPS=12
run pg_w12_r4 TXG_STAGE_PG=1 TXG_STAGE_ROUNDS=4 TXG_STAGE_WARPS=12
run pg_w6_clc_r1 TXG_STAGE_PG=1 TXG_STAGE_CLC=1 TXG_STAGE_ROUNDS=1 TXG_STAGE_WARPS=6
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__m_xbar2l1tex_read_sectors.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__pcsamp_warps_issue_stalled_long_scoreboard,smsp__pcsamp_sample_buffer_full
env TXG_STAGE_PG=1 TXG_STAGE_CLC=1 TXG_STAGE_ROUNDS=1 TXG_STAGE_WARPS=12 timeout 250 ncu --metrics $M --clock-control none -k regex:k_step_stage -s 4 -c 1 --csv --log-file gpurun_out/r2ah_counters_pg_w12_clc.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2> gpurun_out/r2ah_ncu.err
tail -2 gpurun_out/r2ah_ncu.err
