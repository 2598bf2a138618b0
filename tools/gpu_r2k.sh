#!/bin/bash
# r2k: one-triangle iteration kernel: GPU tests, bench with both kernels, ncu capture of the new kernel
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2k_pytest.log
tail -n 5 $O/r2k_pytest.log
B200ADMM_PATH_PROF=1 timeout 600 python bench.py --no-e2e --no-cpu > $O/r2k_bench_tri.json 2> $O/r2k_bench_tri.err
echo "tri rc=$?"
B200ADMM_TALL_TRI=0 timeout 600 python bench.py --no-e2e --no-cpu --no-parity > $O/r2k_bench_full.json 2> $O/r2k_bench_full.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tall_path_tri -c 1 -f -o $O/r2k_tall_path_tri \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --no-parity > $O/r2k_ncu_tri.log 2>&1
python - <<'P'
import json
for f in ("tri", "full"):
    try:
        d = json.loads(open("gpurun_out/r2k_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["phase_s"]["iterate"], d["roofline_iteration"]["us_per_iteration"], d["niter_path"], d.get("parity", {}).get("ok"))
    except Exception as e:
        print(f, e)
P
tail -n 3 $O/r2k_bench_tri.err
