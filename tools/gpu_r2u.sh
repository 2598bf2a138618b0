#!/bin/bash
# r2u (1 GPU): final single-GPU evidence of round 2: GPU tests, smoke, headline bench (both arms), launch list, C3 / C4 configs
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2u_env.txt 2>&1; nproc >> $O/r2u_env.txt
( time timeout 1200 python -m pytest tests -m gpu -q ) > $O/r2u_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2u_pytest.log
tail -n 3 $O/r2u_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2u_smoke.log 2>&1; tail -n 1 $O/r2u_smoke.log
timeout 600 python bench.py --impl reference > $O/r2u_bench_reference.json 2> $O/r2u_bench_reference.err
echo "reference rc=$?"
timeout 900 python bench.py > $O/r2u_bench.json 2> $O/r2u_bench.err
echo "bench rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r2u_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > $O/r2u_launches.log 2>&1
for C in enet wide lad bp; do
  timeout 900 python bench.py --config $C > $O/r2u_config_$C.json 2> $O/r2u_config_$C.err
  echo "$C rc=$?"
done
python - <<'P'
import json
for f in ("r2u_bench", "r2u_bench_reference", "r2u_config_enet", "r2u_config_wide", "r2u_config_lad", "r2u_config_bp"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), round(d["ms_per_step"], 1), (d.get("e2e") or {}).get("value"), (d.get("parity") or {}).get("ok"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, e)
P
