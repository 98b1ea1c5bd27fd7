#!/bin/bash
# usage (one 2-GPU gpurun call): gpurun --gpus 2 --timeout 1500 -- tools/gpu_r2_parity2.sh
# Multi-rank parity on hardware: z-slab runs against the oracle and bit for bit against one rank; logs under gpurun_out/.
mkdir -p gpurun_out
export TXG_MG_LOG=$PWD/gpurun_out/r2_parity_mg_results.jsonl
nvidia-smi -L > gpurun_out/r2_parity_mg.log
( time timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_zz_multi_gpu_bcs.py -m gpu -v --tb=short -p no:cacheprovider ) >> gpurun_out/r2_parity_mg.log 2>&1
tail -25 gpurun_out/r2_parity_mg.log
( time TXG_RUN_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zzz_experimental_lag.py -m gpu -v --tb=short -p no:cacheprovider -k "two_ranks" ) > gpurun_out/r2_parity_lag2.log 2>&1
tail -15 gpurun_out/r2_parity_lag2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-e2e --scaling strong > gpurun_out/r2_bench_n2_strong.json 2> gpurun_out/r2_bench_n2_strong.err; echo "rc=$?"; tail -3 gpurun_out/r2_bench_n2_strong.err; cat gpurun_out/r2_bench_n2_strong.json
