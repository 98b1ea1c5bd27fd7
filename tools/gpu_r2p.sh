#!/bin/bash
# usage (one 2-GPU gpurun call): gpurun --gpus 2 --timeout 1500 -- tools/gpu_r2p.sh
# Multi-rank parity on hardware (z-slabs vs the oracle and bit for bit vs one rank; face BCs; free-slip walls across the slab
# face), then the per-GPU work of 8-GPU strong scaling on 2 GPUs (a 512 x 512 x 128 box: 64 planes per rank) and the N = 2 bench line.
mkdir -p gpurun_out /tmp/txg_cache
export TXG_CASE_CACHE=/tmp/txg_cache
export TXG_MG_LOG=$PWD/gpurun_out/r2p_parity_mg_results.jsonl
rm -f $TXG_MG_LOG
nvidia-smi -L > gpurun_out/r2p_parity_mg.log
# the single-GPU suite first: every test steps through txg_step(n >= 2), i.e. through the two-step CUDA graphs
( time TXG_ASSUME_GPU=1 timeout 900 python -m pytest tests -q -m "gpu and not slow" --tb=short -p no:cacheprovider ) > gpurun_out/r2p_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2p_pytest_gpu.log
( time timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_zz_multi_gpu_bcs.py -m gpu -v --tb=short -p no:cacheprovider ) >> gpurun_out/r2p_parity_mg.log 2>&1
tail -20 gpurun_out/r2p_parity_mg.log
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
timeout 600 python bench.py --size 512 --nz 64 --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2p_n1_64planes.json 2> gpurun_out/r2p_n1_64planes.err
tr 29721 --size 512 --nz 128 --scaling strong --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2p_n2_strong_64planes.json 2> gpurun_out/r2p_n2_strong_64planes.err
TXG_GRAPH=0 tr 29723 --size 512 --nz 128 --scaling strong --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2p_n2_strong_64planes_nograph.json 2> gpurun_out/r2p_n2_strong_64planes_nograph.err
tr 29722 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2p_n2_weak.json 2> gpurun_out/r2p_n2_weak.err
python - <<'PY'
import json
for n in ("n1_64planes", "n2_strong_64planes", "n2_strong_64planes_nograph", "n2_weak"):
    try:
        d = json.load(open("gpurun_out/r2p_%s.json" % n))
        k = {a: round(v["ms"] / max(v["launches"], 1), 3) for a, v in d["kernels"].items() if v["launches"]}
        print(n, "MLUPS %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), k, "with events %.3f host %.3f" % (d["ms_per_step_with_kernel_events"], d["host_enqueue_ms_per_step"]), "launches", d["gpu_launches"], "e2e", d["e2e"] and round(d["e2e"]["value"]), "strong", d.get("strong"))
    except Exception as e:
        print(n, "failed", e)
PY
