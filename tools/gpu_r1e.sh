#!/bin/bash
# Round-1e evidence run (one B200): GPU tests, headline bench, launch list and ncu captures of the fp16 Gram kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $O/r1e_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r1e_pytest.log
timeout 600 python bench.py > $O/r1e_bench.json 2> $O/r1e_bench.err
echo "bench rc=$?" >> $O/r1e_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r1e_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > $O/r1e_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram_pair_h -c 1 -f -o $O/r1e_gram_pair_h \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --n 250016 > $O/r1e_ncu_gram.log 2>&1
timeout 200 python tools/gram_check.py 1000000 10000 3 > $O/r1e_gram_full.log 2>&1
ls -la $O | tail -12
