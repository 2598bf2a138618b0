#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -s -k "readme_benchmark" ) > $O/r2H_pytest.log 2>&1
echo "pytest rc=$?"; grep "readme\]" $O/r2H_pytest.log; tail -n 2 $O/r2H_pytest.log
