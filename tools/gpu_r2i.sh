#!/bin/bash
# r2i: LAD / BP full-size trace forensics, factorisation accuracy, ncu capture of the slice-major Gram kernel
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tests/tools/debug_lad_trace.py lad 100000 1000 > $O/r2i_lad_small.log 2>&1
timeout 400 python tests/tools/debug_lad_trace.py lad 500000 5000 > $O/r2i_lad_full.log 2>&1
timeout 400 python tests/tools/debug_lad_trace.py bp 5000 500000 12 > $O/r2i_bp_full.log 2>&1
timeout 300 python tools/debug_spd_accuracy.py > $O/r2i_spd_accuracy.log 2>&1
timeout 300 python tools/time_factor.py 10000 > $O/r2i_time_factor.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram_pair_h -c 1 -f -o $O/r2i_gram_pair_h \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > $O/r2i_ncu_gram.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r2i_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > $O/r2i_launches.log 2>&1
tail -3 $O/r2i_lad_small.log $O/r2i_lad_full.log $O/r2i_bp_full.log $O/r2i_spd_accuracy.log $O/r2i_time_factor.log
