#!/bin/bash
# r2B (1 GPU): final C3 record (full line with e2e and the CPU sample) after the fused / batched / screened wide steps
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --config wide > $O/r2B_config_wide.json 2> $O/r2B_config_wide.err
echo "wide rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2B_config_wide.json").read().strip().splitlines()[-1])
print(d["value"], d["phase_s"], d["parity"].get("ok"), d["niter_path"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["path_wall_s"], d["cpu_baseline"]["value"])
P
