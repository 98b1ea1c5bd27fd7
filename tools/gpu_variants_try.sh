#!/bin/bash
# usage (one gpurun call, after tools/build_variants.sh here): tools/gpu_variants_try.sh
# One parity test per variant library (porous MRT case against the oracle), then kernel times at 512^3; with and without
# the one-pass step.  JSON lines under gpurun_out/var_*.json.
mkdir -p gpurun_out
export TXG_ASSUME_GPU=1
run() { # name lib env...
  n=$1; lib=$2; shift; shift
  export TAXILA_GPU_LIB=$PWD/taxila-lbm_b200/$lib
  env "$@" timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "porous_mrt_minerals_body or node_class" -p no:cacheprovider 2>&1 | tail -1
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err || tail -3 gpurun_out/var_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open("gpurun_out/var_%s.json"%sys.argv[1]))
k=d["kernels"]
print(sys.argv[1], "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), {n:round(v["ms"]/max(v["launches"],1),3) for n,v in k.items() if v["launches"]}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
}
run default libtaxila_gpu.so
for v in ldna stcs stcg ldna_stcs; do
  [ -f taxila-lbm_b200/libtaxila_gpu_$v.so ] && run $v libtaxila_gpu_$v.so
done
run lag_default libtaxila_gpu.so TXG_LAG=1
run tile192 libtaxila_gpu.so TXG_RHOTILE=1
[ -f taxila-lbm_b200/libtaxila_gpu_cap128.so ] && run tile128 libtaxila_gpu_cap128.so TXG_RHOTILE=1
[ -f taxila-lbm_b200/libtaxila_gpu_cap256.so ] && run tile256 libtaxila_gpu_cap256.so TXG_RHOTILE=1
[ -f taxila-lbm_b200/libtaxila_gpu_ldna.so ] && run lag_ldna libtaxila_gpu_ldna.so TXG_LAG=1
[ -f taxila-lbm_b200/libtaxila_gpu_laghints.so ] && run lag_hints libtaxila_gpu_laghints.so TXG_LAG=1
[ -f taxila-lbm_b200/libtaxila_gpu_laghints.so ] && run lag_hints_r128 libtaxila_gpu_laghints.so TXG_LAG=1 TXG_LAG_ROWS=128 TXG_LAG_PLANES=2
