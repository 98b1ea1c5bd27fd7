#!/bin/bash
# usage: gpurun --gpus 8 --timeout 1200 -- tools/gpu_r2t.sh   (4 ranks: parity incl. uneven slabs and face BCs, then the default bench line at N = 4)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_CASE_CACHE=/tmp/txg_cache
export TXG_MG_LOG=$PWD/gpurun_out/r2t_parity_mg8_results.jsonl
rm -f $TXG_MG_LOG
nvidia-smi -L > gpurun_out/r2t_parity_mg8.log
( time timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_zz_multi_gpu_bcs.py tests/test_gpu_parity.py tests/test_zgpu_face_bcs.py -m gpu -v --tb=short -p no:cacheprovider -k "eight_ranks" ) >> gpurun_out/r2t_parity_mg8.log 2>&1
tail -12 gpurun_out/r2t_parity_mg8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2t_n8_weak.json 2> gpurun_out/r2t_n8_weak.err
tail -2 gpurun_out/r2t_n8_weak.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2t_n8_weak.json"))
k = {a: round(v["ms"] / max(v["launches"], 1), 3) for a, v in d["kernels"].items() if v["launches"]}
print("n8_weak MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), k, "e2e", d["e2e"] and round(d["e2e"]["value"]), d["e2e"]["breakdown_ms"], "strong", d.get("strong"))
PY
