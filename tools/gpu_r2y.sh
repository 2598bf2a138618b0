#!/bin/bash
# r2y (2 GPUs): consensus at the per-rank shape of BASELINE config 5 after the gemv_t change
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 \
    --config consensus --rows 250000 --cols 80000 > $O/r2y_consensus_2gpu.json 2> $O/r2y_consensus_2gpu.err
echo "consensus rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2y_consensus_2gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["phase_s"], d["parity"], d["niter"], d["ms_per_iteration"], d["roofline"]["frac"])
P
