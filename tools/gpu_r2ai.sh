#!/bin/bash
# usage: gpurun --timeout 700 -- tools/gpu_r2ai.sh   (needs profiles/r2ai_park_form.patch applied: TXG_STAGE_PARK=1, 96 registers, 20 warps per SM; tests, A/B, counters)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
timeout 500 python -m pytest tests/test_zgpu_step_forms.py -q -m gpu --tb=short -p no:cacheprovider -k "forms_bit_identical or straddle or escape or gather_prefetch" 2>&1 | tail -8
run() { # name env...
  n=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2ai_$n.json 2> gpurun_out/r2ai_$n.err || tail -3 gpurun_out/r2ai_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2ai_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"] and n.startswith("k_")}, d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"), d["clocks"]["reasons"], "drift", d.get("mass_drift_rel"))
PY
}
run park1 TXG_STAGE_PARK=1
run base1
run park2 TXG_STAGE_PARK=1
run base2
run park_r3 TXG_STAGE_PARK=1 TXG_STAGE_ROUNDS=3
run park_r1 TXG_STAGE_PARK=1 TXG_STAGE_ROUNDS=1
run park_pg_r3 TXG_STAGE_PARK=1 TXG_STAGE_ROUNDS=3 TXG_STAGE_PG=1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__pcsamp_warps_issue_stalled_long_scoreboard,smsp__pcsamp_warps_issue_stalled_short_scoreboard,smsp__pcsamp_warps_issue_stalled_mio_throttle,smsp__pcsamp_warps_issue_stalled_wait,smsp__pcsamp_warps_issue_stalled_not_selected,smsp__pcsamp_warps_issue_stalled_selected,smsp__pcsamp_warps_issue_stalled_math_pipe_throttle,smsp__pcsamp_warps_issue_stalled_lg_throttle
env TXG_STAGE_PARK=1 timeout 250 ncu --metrics $M --clock-control none -k regex:k_step_stage -s 4 -c 1 --csv --log-file gpurun_out/r2ai_counters_park.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2> gpurun_out/r2ai_ncu.err
tail -2 gpurun_out/r2ai_ncu.err
