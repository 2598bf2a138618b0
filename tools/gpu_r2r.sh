#!/bin/bash
# r2r: float64 tensor-core GEMM: kernel tests, isolated timing, LAD / BP configs
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_parity_midsize.py -m gpu -q -k "gemm_f64 or lad or bp" ) > $O/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2r_pytest.log
tail -n 3 $O/r2r_pytest.log
timeout 600 python tools/time_gemm_f64.py > $O/r2r_time_gemm_f64.log 2>&1
cat $O/r2r_time_gemm_f64.log
timeout 400 python bench.py --config lad --no-e2e > $O/r2r_config_lad.json 2> $O/r2r_config_lad.err
echo "lad rc=$?"
timeout 400 python bench.py --config bp --no-e2e > $O/r2r_config_bp.json 2> $O/r2r_config_bp.err
echo "bp rc=$?"
python - <<'P'
import json
for f in ("lad", "bp"):
    try:
        d = json.loads(open("gpurun_out/r2r_config_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["phase_s"], d["parity"].get("ok"), d["niter"], d["setup_flops"])
    except Exception as e:
        print(f, e)
P
