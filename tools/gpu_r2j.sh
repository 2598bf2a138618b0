#!/bin/bash
# r2j: Gram producer lock step (B200ADMM_GRAM_LEAD sweep), ncu capture of the default, GPU tests, LAD / BP configs
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2j_pytest.log
for L in 0 32 8 16 64 128; do
  B200ADMM_GRAM_LEAD=$L timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-parity > $O/r2j_bench_lead$L.json 2> $O/r2j_bench_lead$L.err
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram_pair_h -c 1 -f -o $O/r2j_gram_pair_h \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > $O/r2j_ncu_gram.log 2>&1
timeout 400 python bench.py --config lad --no-e2e > $O/r2j_config_lad.json 2> $O/r2j_config_lad.err
echo "lad rc=$?"
timeout 400 python bench.py --config bp --no-e2e > $O/r2j_config_bp.json 2> $O/r2j_config_bp.err
echo "bp rc=$?"
tail -n 3 $O/r2j_pytest.log
python - <<'P'
import json
for L in (0, 8, 16, 32, 64, 128):
    try:
        d = json.loads(open("gpurun_out/r2j_bench_lead%d.json" % L).read().strip().splitlines()[-1])
        print(L, d["value"], d["phase_s"]["gram"], d["roofline"]["kernel_s"], d["clocks"]["sm_mhz"], d["clocks"]["power_w_median"])
    except Exception as e:
        print(L, e)
P
