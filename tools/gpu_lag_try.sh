#!/bin/bash
# usage (one gpurun call, about 35 GPU-minutes): gpurun --timeout 2400 -- tools/gpu_lag_try.sh
# The opt-in one-pass step (TXG_LAG=1): parity first (tests/test_zzz_experimental_lag.py), then kernel times of the default
# step and of a few (band rows, lag planes, M block size) settings at 512^3, each as one bench.py JSON line under gpurun_out/.
mkdir -p gpurun_out
export TXG_ASSUME_GPU=1
( time TXG_RUN_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_zzz_experimental_lag.py -x -q -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/lag_tests.log 2>&1
tail -15 gpurun_out/lag_tests.log
run() { # name env...
  n=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/lag_$n.json 2> gpurun_out/lag_$n.err || tail -3 gpurun_out/lag_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/lag_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run default
run rhotile TXG_RHOTILE=1
run r64_l2 TXG_LAG=1 TXG_LAG_ROWS=64 TXG_LAG_PLANES=2     # the default setting
run r64_l2_tile TXG_LAG=1 TXG_RHOTILE=1 TXG_LAG_ROWS=64 TXG_LAG_PLANES=2   # one-pass step + density tiles
run r64_l1 TXG_LAG=1 TXG_LAG_ROWS=64 TXG_LAG_PLANES=1
run r64_l3 TXG_LAG=1 TXG_LAG_ROWS=64 TXG_LAG_PLANES=3
run r64_l0 TXG_LAG=1 TXG_LAG_ROWS=64 TXG_LAG_PLANES=0
run r128_l1 TXG_LAG=1 TXG_LAG_ROWS=128 TXG_LAG_PLANES=1
run r128_l2 TXG_LAG=1 TXG_LAG_ROWS=128 TXG_LAG_PLANES=2
run r256_l1 TXG_LAG=1 TXG_LAG_ROWS=256 TXG_LAG_PLANES=1
run r64_l2_m2048 TXG_LAG=1 TXG_LAG_ROWS=64 TXG_LAG_PLANES=2 TXG_LAG_MPOS=2048
run r64_l2_m128 TXG_LAG=1 TXG_LAG_ROWS=64 TXG_LAG_PLANES=2 TXG_LAG_MPOS=128
# DRAM traffic of the one-pass kernel (the question: do the density reads hit L2?) -- one launch under ncu
TXG_LAG=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_step_fused_lag -s 4 -c 1 --csv --log-file gpurun_out/lag_ncu.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2> gpurun_out/lag_ncu.err
tail -3 gpurun_out/lag_ncu.csv
