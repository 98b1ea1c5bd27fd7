#!/bin/bash
# usage (one gpurun call, ~25 GPU-minutes): gpurun --timeout 2100 -- tools/gpu_r2c.sh
# Band blocks (k_step_band, default) on a B200: the whole -m gpu suite, oracle parity at 128^3 x 1000 / 128^3 / 256^3
# (tests/test_zgpu_large_parity.py), then 512^3 timings against the table-driven kernel (TXG_BAND=0) and one ncu pass.
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
nvidia-smi -L > gpurun_out/r2c_pytest_gpu.log
( time timeout 900 python -m pytest tests -x -q -m "gpu and not slow" --tb=short -p no:cacheprovider ) >> gpurun_out/r2c_pytest_gpu.log 2>&1
tail -8 gpurun_out/r2c_pytest_gpu.log
rm -f gpurun_out/r2c_parity_large.jsonl
( time TXG_PARITY_LOG=$PWD/gpurun_out/r2c_parity_large.jsonl timeout 1200 python -m pytest tests/test_zgpu_large_parity.py -v -m gpu --tb=short -p no:cacheprovider ) > gpurun_out/r2c_parity_large.log 2>&1
tail -8 gpurun_out/r2c_parity_large.log
run() { # name env...
  n=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2c_$n.json 2> gpurun_out/r2c_$n.err || tail -3 gpurun_out/r2c_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2c_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run band_lb1024
run table TXG_BAND=0
run band_lb512 TXG_BAND_LB=512
run band_lb2048 TXG_BAND_LB=2048
run band_lb1024_pf0 TXG_BAND_PF=0
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:k_step_band -s 4 -c 1 --csv --log-file gpurun_out/r2c_band_ncu.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2> gpurun_out/r2c_band_ncu.err
tail -12 gpurun_out/r2c_band_ncu.csv
