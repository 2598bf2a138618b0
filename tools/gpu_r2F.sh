#!/bin/bash
# r2F (1 GPU): ncu captures of the new round-2 kernels outside the headline path (wide screen, fused active-set step, f64 DMMA GEMM)
set -u
O=gpurun_out; mkdir -p $O
timeout 500 ncu --set full --clock-control none -k regex:wide_screen -s 20 -c 1 -f -o $O/r2F_wide_screen \
    python bench.py --config wide --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > $O/r2F_ncu_wide_screen.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:wide_active_ax -s 3000 -c 1 -f -o $O/r2F_wide_active_ax \
    python bench.py --config wide --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity --nlambda 60 > $O/r2F_ncu_wide_active_ax.log 2>&1
timeout 500 ncu --set full --clock-control none -k regex:gemm_f64_mma -c 1 -f -o $O/r2F_gemm_f64_mma \
    python bench.py --config lad --maxit 3 --no-e2e --no-cpu --no-parity > $O/r2F_ncu_gemm_f64.log 2>&1
ls -la $O | grep r2F
