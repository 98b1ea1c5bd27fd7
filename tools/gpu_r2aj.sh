#!/bin/bash
# usage: gpurun --timeout 400 -- tools/gpu_r2aj.sh   (where the e2e time outside the steps goes: device-synchronised marks)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
TXG_BENCH_SYNC_MARKS=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2aj_sync.json 2> gpurun_out/r2aj_sync.err || tail -3 gpurun_out/r2aj_sync.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2aj_sync.json"))
print("e2e %.0f" % d["e2e"]["value"], d["e2e"]["breakdown_ms"], d["ms_per_step"])
PY
