#!/bin/bash
# Round-2 evidence on one B200 (gpurun --timeout 2400 -- tools/gpu_r2_evidence.sh): the whole -m gpu suite incl. the large-size
# oracle comparisons, the default bench line, the CPU arm, the ncu launch list of the same command, one full ncu capture per hot kernel
# (SKIP_NCU_FULL=1 leaves the two full captures out; R=name sets the prefix of the files).
O=gpurun_out; R=${R:-r2zz}
mkdir -p $O /tmp/txg_cache
export TXG_CASE_CACHE=/tmp/txg_cache
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/${R}_nvsmi.csv 2>&1
rm -f $O/${R}_parity_large.jsonl
( time TXG_PARITY_LOG=$PWD/$O/${R}_parity_large.jsonl timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider ) > $O/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${R}_pytest_gpu.log
timeout 900 python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err; echo "bench rc=$?"; cat $O/${R}_bench.json | cut -c1-1500
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $O/${R}_bench_reference.json 2> $O/${R}_bench_reference.err; echo "bench ref rc=$?"; cat $O/${R}_bench_reference.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/${R}_ncu_launch.log 2>&1; echo "ncu list rc=$?"
[ -n "$SKIP_NCU_FULL" ] || timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step_stage -s 3 -c 1 -o $O/${R}_step_stage -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/${R}_ncu_step_stage.log 2>&1; echo "ncu stage rc=$?"
[ -n "$SKIP_NCU_FULL" ] || timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_moments -s 3 -c 1 -o $O/${R}_moments -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $O/${R}_ncu_moments.log 2>&1; echo "ncu moments rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/${R}_smoke.log
