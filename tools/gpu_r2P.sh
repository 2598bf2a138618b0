#!/bin/bash
# r2P (1 GPU): launch list (duration, grid) of the factorisation a <- a^-1 at p = 1e4, kernel by kernel (graph replay off)
set -u
O=gpurun_out; mkdir -p $O
B200ADMM_GRAPH=0 timeout 170 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -c 2600 --csv \
    --log-file $O/r2P_factor_launches.csv python tools/time_factor.py 10000 fast > $O/r2P_time_factor.log 2>&1
tail -n 2 $O/r2P_time_factor.log; wc -l $O/r2P_factor_launches.csv
