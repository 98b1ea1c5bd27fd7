#!/bin/bash
# First GPU pass of the round: parity tests, bench line, ncu launch list, one full capture per hot kernel.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi > $O/nvsmi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 900 python bench.py --steps 30 --warmup 5 > $O/bench_pipe.json 2> $O/bench_pipe.err; echo "bench rc=$?"
cat $O/bench_pipe.json
TXG_NO_PIPE=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu > $O/bench_nopipe.json 2> $O/bench_nopipe.err; echo "bench nopipe rc=$?"
cat $O/bench_nopipe.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_512.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_launch.log 2>&1; echo "ncu list rc=$?"
TXG_NO_PIPE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 3 -c 1 -o $O/collide_full_512 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_collide.log 2>&1; echo "ncu collide rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_collide -s 3 -c 1 -o $O/collide_pipe_full_512 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_collide_pipe.log 2>&1; echo "ncu collide pipe rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_moments -s 3 -c 1 -o $O/moments_full_512 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/ncu_moments.log 2>&1; echo "ncu moments rc=$?"
ls -la $O
