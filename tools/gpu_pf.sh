#!/bin/bash
# prefetch-distance sweep of the collide kernel (512^3 porous, kernel times from bench.py)
mkdir -p gpurun_out
for pf in 0 148 296 592 1184 2368; do
  TXG_PF=$pf timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/pf_$pf.json 2> gpurun_out/pf_$pf.err || tail -3 gpurun_out/pf_$pf.err
  python - $pf <<'PY'
import json,sys
d=json.load(open("gpurun_out/pf_%s.json"%sys.argv[1]))
k=d["kernels"]
print("PF",sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]}, d["clocks"])
PY
done
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
