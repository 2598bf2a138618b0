#!/bin/bash
# r2N (1 GPU): regression at the last product commit of the round: all GPU tests, smoke, headline bench
set -u
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests -m gpu -q -x ) > $O/r2N_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2N_pytest.log; grep -v "^$" $O/r2N_pytest.log | tail -n 6
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2N_smoke.log 2>&1; tail -n 1 $O/r2N_smoke.log
( time timeout 400 python bench.py > $O/r2N_bench.json 2> $O/r2N_bench.err ) 2>&1 | grep real
echo "bench rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2N_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["phase_s"], d["e2e"]["value"], d["parity"]["ok"], d["clocks"]["sm_mhz"], d["roofline"]["frac"])
P
