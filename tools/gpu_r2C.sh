#!/bin/bash
# r2C (2 GPUs): device-controlled consensus batches: 1-GPU tests (bit identity with the host loop, oracle parity), multi-GPU check,
# consensus at the per-rank shape of config 5
set -u
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py tests/test_gpu_multi.py -m gpu -q -k "consensus or parallel or gemv or multi or sharded" ) > $O/r2C_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 $O/r2C_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 \
    --config consensus --rows 250000 --cols 80000 > $O/r2C_consensus_2gpu.json 2> $O/r2C_consensus_2gpu.err
echo "consensus rc=$?"
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2C_consensus_2gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["phase_s"], d["parity"]["ok"], d["niter"], d["ms_per_iteration"], d["roofline"]["frac"])
P
