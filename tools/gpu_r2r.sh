#!/bin/bash
# usage: gpurun --timeout 900 -- tools/gpu_r2r.sh   (the default bench line with the e2e breakdown; CPU arm)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_CASE_CACHE=/tmp/txg_cache
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2r_bench20.json 2> gpurun_out/r2r_bench20.err; tail -2 gpurun_out/r2r_bench20.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2r_bench20.json"))
print("MLUPS %.0f ms/step %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["e2e"]["breakdown_ms"], d["roofline"]["frac"], d["step_roofline"]["frac_of_hbm_peak"])
PY
