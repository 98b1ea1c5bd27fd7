#!/bin/bash
# usage: gpurun --timeout 400 -- tools/gpu_r2al.sh   (the driver's N = 1 command twice, without the CPU leg: e2e and its breakdown)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_CASE_CACHE=/tmp/txg_cache
for i in 1 2; do
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2al_bench20_$i.json 2> gpurun_out/r2al_bench20_$i.err || tail -3 gpurun_out/r2al_bench20_$i.err
done
python - <<'PY'
import json
for n in (1, 2):
    d=json.load(open("gpurun_out/r2al_bench20_%d.json" % n))
    print(n, "value %.0f ms/step %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["e2e"]["breakdown_ms"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
