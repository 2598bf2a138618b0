"""Row-split consensus lasso (the reference's $parallel(N), BASELINE.json config 5) with one block per GPU and
one NCCL all-reduce per iteration.  Launch with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_consensus.py [--rows 1000000] [--cols 80000] [--lambda-frac 0.1] [--maxit 2000]

Each rank generates its own row block of the synthetic design in HBM (counter-based generator: the same
matrix whatever N).  Prints one JSON line on rank 0: setup (DataStd, block Gram + inverse), iterations,
seconds per iteration against the HBM bound of the K_i^-1 product, and the all-reduce payload."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import admm_b200
from admm_b200 import _capi as K
from admm_b200 import dist as D


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", dest="n", type=int, default=1_000_000)
    ap.add_argument("--cols", dest="p", type=int, default=80_000)
    ap.add_argument("--lambda-frac", type=float, default=0.1, help="lambda as a fraction of lambda_max")
    ap.add_argument("--maxit", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=123)
    a = ap.parse_args()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    r0, nr = D.row_block(a.n, world, rank)
    X = torch.empty((a.p, nr), dtype=torch.float32, device="cuda")
    y = torch.empty(nr, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_synth_f32(X.data_ptr(), y.data_ptr(), nr, a.p, r0, a.seed, 0.0, 2.0, min(100, a.p), 1.0))
    D.init_comm()
    # lambda_max from a two-point path at maxit = 1 (the grid's first value is lambda_max)
    f0 = admm_b200.admm_lasso(X.t(), y).penalty(nlambda=2).parallel(world).opts(maxit=1).fit()
    lam = [a.lambda_frac * float(f0.lambda_[0])]
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    f = admm_b200.admm_lasso(X.t(), y).penalty(lam).parallel(world).opts(maxit=a.maxit).fit()
    torch.cuda.synchronize(); dist.barrier()
    wall = time.perf_counter() - t0
    T = f.info["timing"]
    it = int(f.niter.sum())
    t = torch.tensor([wall, T["gram"], T["iterate"], T["standardize"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, t_setup, t_iter, t_std = (float(v) for v in t)
    if rank == 0:
        bpi = 4.0 * a.p * (a.p + 1) + 64.0 * a.p
        try:
            peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
        except Exception:
            peak = 6545.0
        print(json.dumps({
            "bench": "consensus_lasso_rowsplit", "n": a.n, "p": a.p, "n_gpus": world, "rows_per_gpu": nr, "lambda": lam[0],
            "niter": it, "converged": bool(it <= a.maxit), "wall_s": wall, "standardize_s": t_std, "block_gram_and_inverse_s": t_setup,
            "iterate_s": t_iter, "ms_per_iter": t_iter / max(it, 1) * 1e3, "bytes_per_iter_per_gpu": bpi,
            "iter_gbs_per_gpu": bpi * it / max(t_iter, 1e-9) / 1e9, "hbm_frac": bpi * it / max(t_iter, 1e-9) / 1e9 / peak,
            "allreduce_payload_bytes": 4 * (a.p + 3), "nnz": int(f.beta.nnz), "rho": f.info["rho"],
            "device": admm_b200.device_info()["name"]}), flush=True)
    D.destroy_comm()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
