#!/bin/bash
# r2l: what bounds the one-triangle sweep?  (timing experiments: results of modes 1 / 2 are numerically meaningless)
set -u
mkdir -p gpurun_out
O=gpurun_out
for M in 0 1 2; do
  B200ADMM_TRI_DBG=$M B200ADMM_PATH_PROF=1 timeout 300 python bench.py --steps 2 --warmup 1 --nlambda 20 --maxit 50 --no-e2e --no-cpu --no-parity > $O/r2l_dbg$M.json 2> $O/r2l_dbg$M.err
  tail -n 1 $O/r2l_dbg$M.err
done
B200ADMM_TALL_TRI=0 B200ADMM_PATH_PROF=1 timeout 300 python bench.py --steps 2 --warmup 1 --nlambda 20 --maxit 50 --no-e2e --no-cpu --no-parity > $O/r2l_full.json 2> $O/r2l_full.err
tail -n 1 $O/r2l_full.err
