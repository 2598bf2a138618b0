#!/bin/bash
# Final round-1 evidence run (one B200): GPU tests (also with the graph / tensor-factor switches off), headline bench
# (both arms), launch list, ncu --set full of the Gram, fused-split and iteration kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $O/r1s_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r1s_pytest.log
( B200ADMM_GRAPH=0 B200ADMM_FACTOR_TENSOR=0 timeout 300 python -m pytest tests/test_gpu_tall.py tests/test_gpu_kernels.py -m gpu -x -q -k "not scale" ) > $O/r1s_pytest_switches_off.log 2>&1
echo "pytest rc=$?" >> $O/r1s_pytest_switches_off.log
timeout 600 python bench.py > $O/r1s_bench.json 2> $O/r1s_bench.err
timeout 300 python bench.py --impl reference > $O/r1s_bench_reference.json 2> $O/r1s_bench_reference.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r1s_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > $O/r1s_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram_pair_h -c 1 -f -o $O/r1s_gram_pair_h \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > $O/r1s_ncu_gram.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:std_split_xty -c 1 -f -o $O/r1s_std_split_xty \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > $O/r1s_ncu_split.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:tall_path -c 1 -f -o $O/r1s_tall_path \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > $O/r1s_ncu_tall.log 2>&1
ls -la $O | tail -14
