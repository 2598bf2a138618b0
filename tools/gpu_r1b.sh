#!/bin/bash
# Round-1b evidence run (one B200): GPU tests, headline bench, launch list and ncu captures.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r1b.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/r1b_smi.txt 2>&1
nproc >> $O/r1b_smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/r1b_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r1b_pytest.log
timeout 600 python bench.py > $O/r1b_bench.json 2> $O/r1b_bench.err
echo "bench rc=$?" >> $O/r1b_bench.err
# launch list of one step (same command, profiler-serialised: compare shares only)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r1b_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > $O/r1b_launches.log 2>&1
# full captures: the Gram kernel (reduced n so 40 replays stay short), the iteration kernel, the fused z/u kernel
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram_pair -c 1 -f -o $O/r1b_gram_pair \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --n 250000 > $O/r1b_ncu_gram.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tall_path -c 1 -f -o $O/r1b_tall_path \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --n 100000 > $O/r1b_ncu_tall.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_zu -c 1 -f -o $O/r1b_fused_zu \
    python tools/bench_configs.py zu --repeats 2 > $O/r1b_ncu_zu.log 2>&1
timeout 120 python tools/bench_configs.py zu > $O/r1b_zu.json 2>&1
ls -la $O
