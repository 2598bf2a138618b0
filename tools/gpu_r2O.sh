#!/bin/bash
# r2O (1 GPU): ncu capture of the factorisation's tensor-core kernel (tn_pair_kernel: rank-1024 trailing update / inverse products)
set -u
O=gpurun_out; mkdir -p $O
B200ADMM_GRAPH=0 timeout 170 ncu --set full --clock-control none -k regex:tn_pair_kernel -s 60 -c 3 -f -o $O/r2O_tn_pair \
    python tools/time_factor.py 10000 fast > $O/r2O_ncu_tn_pair.log 2>&1
tail -n 4 $O/r2O_ncu_tn_pair.log; ls -la $O | grep r2O
