#!/bin/bash
# usage: gpurun --gpus 2 --timeout 1200 -- tools/gpu_r2q.sh   (comm stream at high priority, packed f halo: multi-rank parity + strong / weak lines)
mkdir -p gpurun_out /tmp/txg_cache
export TXG_CASE_CACHE=/tmp/txg_cache
export TXG_MG_LOG=$PWD/gpurun_out/r2q_parity_mg_results.jsonl
rm -f $TXG_MG_LOG
nvidia-smi -L > gpurun_out/r2q_parity_mg.log
( time timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_zz_multi_gpu_bcs.py -m gpu -v --tb=short -p no:cacheprovider ) >> gpurun_out/r2q_parity_mg.log 2>&1
tail -14 gpurun_out/r2q_parity_mg.log
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
tr 29721 --size 512 --nz 128 --scaling strong --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2q_n2_strong_64planes.json 2> gpurun_out/r2q_n2_strong_64planes.err
tr 29722 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2q_n2_weak.json 2> gpurun_out/r2q_n2_weak.err
python - <<'PY'
import json
for n in ("n2_strong_64planes", "n2_weak"):
    try:
        d = json.load(open("gpurun_out/r2q_%s.json" % n))
        k = {a: round(v["ms"] / max(v["launches"], 1), 3) for a, v in d["kernels"].items() if v["launches"]}
        print(n, "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), k, "with events %.3f host %.3f" % (d["ms_per_step_with_kernel_events"], d["host_enqueue_ms_per_step"]), "launches", d["gpu_launches"], "e2e", d["e2e"] and round(d["e2e"]["value"]), "strong", d.get("strong"))
    except Exception as e:
        print(n, "failed", e)
PY
