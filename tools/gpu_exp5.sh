#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { # name env...
  n=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/x_$n.json 2> gpurun_out/x_$n.err || tail -3 gpurun_out/x_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/x_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]})
PY
}
for o in 0 1 2 4 6 7; do run stream_o$o TXG_STREAM=1 TXG_OPTS=$o; done
TXG_OPTS=7 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 3 -c 1 -o gpurun_out/collide_stream7_512 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_collide_stream7.log 2>&1; echo "ncu rc=$?"
