#!/bin/bash
# r2o (8 GPUs): BASELINE config 5 (consensus n = 1e6 x p = 8e4, one row block per GPU), multi-GPU check, tall bench
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2o_env.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 \
    --config consensus > $O/r2o_consensus_8gpu.json 2> $O/r2o_consensus_8gpu.err
echo "consensus 8gpu rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tests/mgpu_check.py \
    > $O/r2o_mgpu_check_8gpu.log 2>&1
echo "mgpu_check 8gpu rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 \
    > $O/r2o_bench_8gpu.json 2> $O/r2o_bench_8gpu.err
echo "tall 8gpu rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 \
    --no-e2e > $O/r2o_bench_4gpu.json 2> $O/r2o_bench_4gpu.err
echo "tall 4gpu rc=$?"
tail -n 2 $O/r2o_mgpu_check_8gpu.log
tail -c 800 $O/r2o_consensus_8gpu.err
python - <<'P'
import json
for f in ("r2o_consensus_8gpu", "r2o_bench_8gpu", "r2o_bench_4gpu"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["phase_s"], d.get("parity_vs_n1", d.get("parity", {})).get("ok"), d.get("niter", d.get("niter_path")), d.get("ms_per_iteration"))
    except Exception as e:
        print(f, e)
P
