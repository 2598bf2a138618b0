#!/bin/bash
# r2L (1 GPU): the README's p > n rows against glmnet restated at its default threshold, through the CUDA library
set -u
O=gpurun_out; mkdir -p $O
( time timeout 400 python -m pytest tests/test_gpu_models.py -m gpu -q -s -k "readme_benchmark_wide_rows" ) > $O/r2L_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2L_pytest.log; grep -v "^$" $O/r2L_pytest.log | tail -n 12
