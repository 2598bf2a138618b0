#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -k "gemm_f64 or lad or bp" ) > $O/r2G_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 2 $O/r2G_pytest.log
timeout 600 python tools/time_gemm_f64.py tensor > $O/r2G_time_gemm_f64.log 2>&1
cat $O/r2G_time_gemm_f64.log
