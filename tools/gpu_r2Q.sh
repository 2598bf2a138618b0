#!/bin/bash
# r2Q (1 GPU): README tall columns through the CUDA library with the rho of the build that knitted the README
set -u
O=gpurun_out; mkdir -p $O
( timeout 50 python -m pytest tests/test_gpu_tall.py -m gpu -q -s -k "print_precision" ) > $O/r2Q_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2Q_pytest.log; grep -v "^$" $O/r2Q_pytest.log | tail -n 8
