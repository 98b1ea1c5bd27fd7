#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n2_weak.json 2> gpurun_out/bench_n2_weak.err; echo "rc=$?"; tail -3 gpurun_out/bench_n2_weak.err; cat gpurun_out/bench_n2_weak.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-e2e --scaling strong > gpurun_out/bench_n2_strong.json 2> gpurun_out/bench_n2_strong.err; echo "rc=$?"; tail -3 gpurun_out/bench_n2_strong.err; cat gpurun_out/bench_n2_strong.json
