#!/bin/bash
# r2I (1 GPU): GPU tests + smoke at the last commit of the round
set -u
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q -s ) > $O/r2I_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2I_pytest.log; tail -n 4 $O/r2I_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2I_smoke.log 2>&1; tail -n 1 $O/r2I_smoke.log
