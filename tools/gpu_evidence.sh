#!/bin/bash
# Round evidence on one B200: parity tests, the bench line, the reference (CPU) arm, the ncu launch list of
# the same command, one full ncu capture per hot kernel.  Outputs in gpurun_out/ (summaries are copied to profiles/).
R=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/${R}_nvsmi.csv 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${R}_pytest_gpu.log
timeout 900 python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err; echo "bench rc=$?"; cat $O/${R}_bench.json
timeout 600 python bench.py --impl reference --steps 3 > $O/${R}_bench_reference.json 2> $O/${R}_bench_reference.err; echo "bench ref rc=$?"; cat $O/${R}_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/${R}_ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_fused -s 3 -c 1 -o $O/${R}_step_fused -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/${R}_ncu_step_fused.log 2>&1; echo "ncu fused rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_moments -s 3 -c 1 -o $O/${R}_moments -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/${R}_ncu_moments.log 2>&1; echo "ncu moments rc=$?"
