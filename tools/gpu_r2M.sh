#!/bin/bash
# r2M (1 GPU): the Rcpp glue executed against the CUDA library (functional <Rcpp.h> stand-in)
set -u
O=gpurun_out; mkdir -p $O
( time timeout 400 python -m pytest tests/test_gpu_rglue.py -m gpu -q -s ) > $O/r2M_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2M_pytest.log; grep -v "^$" $O/r2M_pytest.log | tail -n 30
