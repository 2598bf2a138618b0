#!/bin/bash
# r2t: warp-tile sweep of the one-triangle kernel vs the thread-column sweep; pinned finish
set -u
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_tall.py tests/test_gpu_parity_midsize.py -m gpu -q -x ) > $O/r2t_pytest_sweep1.log 2>&1
echo "pytest(sweep 1) rc=$?" >> $O/r2t_pytest_sweep1.log
tail -n 3 $O/r2t_pytest_sweep1.log
( B200ADMM_TRI_SWEEP=0 timeout 900 python -m pytest tests/test_gpu_tall.py tests/test_gpu_parity_midsize.py -m gpu -q -x -k "tall or triangle or exhausted" ) > $O/r2t_pytest_sweep0.log 2>&1
echo "pytest(sweep 0) rc=$?" >> $O/r2t_pytest_sweep0.log
tail -n 3 $O/r2t_pytest_sweep0.log
for M in 1 0; do
  B200ADMM_TRI_SWEEP=$M B200ADMM_PATH_PROF=1 timeout 600 python bench.py --no-e2e --no-cpu > $O/r2t_bench_sweep$M.json 2> $O/r2t_bench_sweep$M.err
  echo "sweep $M rc=$?"; tail -n 1 $O/r2t_bench_sweep$M.err
done
python - <<'P'
import json
for f in ("1", "0"):
    try:
        d = json.loads(open("gpurun_out/r2t_bench_sweep%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["phase_s"]["iterate"], d["phase_s"]["finish"], d["roofline_iteration"]["us_per_iteration"], d["niter_path"], d.get("parity", {}).get("ok"), d["parity"]["max_abs_dbeta"])
    except Exception as e:
        print(f, e)
P
