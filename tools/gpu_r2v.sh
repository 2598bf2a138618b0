#!/bin/bash
# r2v (8 GPUs): final multi-GPU evidence: multi-GPU check on 8 ranks, tall bench at N = 8, 4, 2 (same box)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py \
    > $O/r2v_mgpu_check_8gpu.log 2>&1
echo "mgpu_check 8gpu rc=$?"
grep "mgpu_check ok" $O/r2v_mgpu_check_8gpu.log
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N \
      > $O/r2v_bench_${N}gpu.json 2> $O/r2v_bench_${N}gpu.err
  echo "tall ${N}gpu rc=$?"
done
python - <<'P'
import json
for N in (8, 4, 2):
    try:
        d = json.loads(open("gpurun_out/r2v_bench_%dgpu.json" % N).read().strip().splitlines()[-1])
        print(N, round(d["value"], 1), {k: round(v, 4) for k, v in d["phase_s"].items()}, d["parity_vs_n1"]["ok"], round(d["e2e"]["value"], 1), d["e2e"].get("max_abs_dbeta_vs_device_input"))
    except Exception as e:
        print(N, e)
P
