#!/bin/bash
# r2z (8 GPUs): BASELINE config 5 after the gemv_t change
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 \
    --config consensus > $O/r2z_consensus_8gpu.json 2> $O/r2z_consensus_8gpu.err
echo "consensus 8gpu rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2z_consensus_8gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["phase_s"], d["parity"]["ok"], d["niter"], d["ms_per_iteration"], d["roofline"]["frac"])
P
