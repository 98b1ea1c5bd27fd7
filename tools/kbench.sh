#!/bin/bash
# usage: tools/kbench.sh <size> <lib1.so> [lib2.so ...]  -- bench.py kernel times per library build
size=$1; shift
for lib in "$@"; do
  TAXILA_GPU_LIB=$PWD/$lib timeout 600 python bench.py --size $size --steps 10 --warmup 3 --no-e2e --no-cpu > /tmp/kb.json 2>/tmp/kb.err || tail -3 /tmp/kb.err
  python - "$lib" <<'PY'
import json,sys
d=json.load(open("/tmp/kb.json"))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]})
PY
done
