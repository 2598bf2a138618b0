#!/bin/bash
# r2q (1 GPU): full GPU tests, headline bench (both arms), launch list, ncu capture of the one-triangle iteration kernel
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > $O/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2q_pytest.log
tail -n 3 $O/r2q_pytest.log
timeout 900 python bench.py > $O/r2q_bench.json 2> $O/r2q_bench.err
echo "bench rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r2q_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > $O/r2q_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tall_path_tri -c 1 -f -o $O/r2q_tall_path_tri \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > $O/r2q_ncu_tri.log 2>&1
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2q_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["phase_s"], d["e2e"]["value"], d["parity"]["ok"], d["roofline_iteration"]["us_per_iteration"], d["cpu_baseline"]["value"])
P
