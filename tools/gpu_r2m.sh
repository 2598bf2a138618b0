#!/bin/bash
# r2m (2 GPUs): GPU tests incl. the multi-GPU check, tall bench on 2 ranks (parity_vs_n1), consensus rehearsal at the
# per-rank shape of BASELINE config 5 (125 000 rows x 80 000 columns per rank)
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2m_env.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2m_pytest.log
tail -n 4 $O/r2m_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
    > $O/r2m_bench_2gpu.json 2> $O/r2m_bench_2gpu.err
echo "tall 2gpu rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 \
    --config consensus --rows 250000 --cols 80000 > $O/r2m_consensus_2gpu_rehearsal.json 2> $O/r2m_consensus_2gpu_rehearsal.err
echo "consensus rehearsal rc=$?"
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> $O/r2m_env.txt
tail -c 600 $O/r2m_bench_2gpu.err; tail -c 1500 $O/r2m_consensus_2gpu_rehearsal.err
python - <<'P'
import json
for f in ("r2m_bench_2gpu", "r2m_consensus_2gpu_rehearsal"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["phase_s"], d.get("parity_vs_n1", d.get("parity", {})).get("ok"), d.get("niter", d.get("niter_path")))
    except Exception as e:
        print(f, e)
P
