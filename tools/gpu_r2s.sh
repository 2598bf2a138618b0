#!/bin/bash
# r2s: screened regular steps of the wide solver: tests, C3 bench with and without the screen; fp64 gemm tests
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_models.py tests/test_gpu_parity_midsize.py tests/test_gpu_kernels.py -m gpu -q -k "wide or gemm_f64 or readme" ) > $O/r2s_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2s_pytest.log
tail -n 3 $O/r2s_pytest.log
timeout 600 python bench.py --config wide --no-cpu > $O/r2s_config_wide.json 2> $O/r2s_config_wide.err
echo "wide rc=$?"
B200ADMM_WIDE_BATCH=0 timeout 600 python bench.py --config wide --no-cpu --no-e2e > $O/r2s_config_wide_unscreened.json 2> $O/r2s_config_wide_unscreened.err
python - <<'P'
import json
for f in ("r2s_config_wide", "r2s_config_wide_unscreened"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["phase_s"], d["parity"].get("ok"), d["niter_path"], d["roofline"].get("mean_columns_evaluated_exactly_per_regular_step"), d["roofline"]["frac"], d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, e)
P
