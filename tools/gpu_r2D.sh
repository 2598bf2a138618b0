#!/bin/bash
# r2D (1 GPU): last regression of the round: GPU tests, smoke, headline bench
set -u
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q ) > $O/r2D_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2D_pytest.log; tail -n 3 $O/r2D_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2D_smoke.log 2>&1; tail -n 1 $O/r2D_smoke.log
timeout 900 python bench.py > $O/r2D_bench.json 2> $O/r2D_bench.err
echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2D_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["phase_s"], d["e2e"]["value"], d["parity"]["ok"], d["clocks"]["sm_mhz"])
P
