#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py::test_output_files_match_reference_golden -m gpu -x -q ) 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29710 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n4_weak.json 2> gpurun_out/bench_n4_weak.err; echo "rc=$?"; tail -2 gpurun_out/bench_n4_weak.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_n4_weak.json')); print('N=4 weak', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
