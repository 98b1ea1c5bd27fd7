#!/bin/bash
# usage (one gpurun call, after tools/build_variants_r2.sh here): gpurun --timeout 900 -- tools/gpu_r2f.sh
# Cache-policy builds of the table-driven K2 (k_step_fused, TXG_STAGE=0) and of the staged K2 at 512^3: one parity test each, then kernel times.
mkdir -p gpurun_out /tmp/txg_cache
export TXG_ASSUME_GPU=1 TXG_CASE_CACHE=/tmp/txg_cache
run() { # name lib env...
  n=$1; lib=$2; shift; shift
  export TAXILA_GPU_LIB=$PWD/taxila-lbm_b200/$lib
  env "$@" timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "porous_mrt_minerals_body or node_class" -p no:cacheprovider 2>&1 | tail -1
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2f_$n.json 2> gpurun_out/r2f_$n.err || tail -3 gpurun_out/r2f_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2f_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run table libtaxila_gpu.so TXG_STAGE=0
for v in ldna stcs stcg ldna_stcs; do
  run table_$v libtaxila_gpu_$v.so TXG_STAGE=0
done
run stage_c2_stcs libtaxila_gpu_stcs.so TXG_STAGE_CHUNKS=2
run stage_c1 libtaxila_gpu.so TXG_STAGE_CHUNKS=1
