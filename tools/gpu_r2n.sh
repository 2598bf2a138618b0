#!/bin/bash
# r2n (2 GPUs): GPU tests incl. the multi-GPU check; consensus rehearsal at the per-rank shape of BASELINE config 5
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > $O/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2n_pytest.log
tail -n 4 $O/r2n_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 \
    --config consensus --rows 250000 --cols 80000 > $O/r2n_consensus_2gpu_rehearsal.json 2> $O/r2n_consensus_2gpu_rehearsal.err
echo "consensus rehearsal rc=$?"
nvidia-smi --query-gpu=memory.used,memory.total --format=csv > $O/r2n_env.txt
tail -c 1500 $O/r2n_consensus_2gpu_rehearsal.err
python - <<'P'
import json
for f in ("r2n_consensus_2gpu_rehearsal",):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["phase_s"], d.get("parity", {}), d.get("niter"), d.get("ms_per_iteration"))
    except Exception as e:
        print(f, e)
P
