"""Multi-GPU parity check, run as `torchrun --nproc-per-node N tests/mgpu_check.py` (one rank per GPU).

1. Row-sharded tall lasso (global DataStd / X'y / Gram by NCCL all-reduce, iterations replicated)
   must reproduce the single-GPU fit of the same matrix: same lambdas, same iteration counts,
   coefficients within float32 summation-order noise (the all-reduce adds the partial Gram matrices
   in a different order than one GPU does).
2. Consensus lasso with one block per rank and one all-reduce per iteration must reproduce the
   CPU oracle's row-split consensus (and the single-process N-block run of the library).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import admm_b200
    from admm_b200 import dist as D
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()

    rng = np.random.default_rng(11)
    n, p = 6001, 160
    x = np.asfortranarray(rng.normal(0.5, 2.0, size=(n, p)))
    b = np.zeros(p); b[:10] = rng.uniform(0.5, 1.5, size=10)
    y = 1.0 + x @ b + rng.normal(size=n)
    r0, nr = D.row_block(n, world, rank)
    xs, ys = np.asfortranarray(x[r0:r0 + nr]), y[r0:r0 + nr].copy()

    # single-GPU references first (no communicator installed yet)
    ref = admm_b200.admm_lasso(x, y).penalty(nlambda=15).fit()
    lam_c = [0.3, 0.1]
    ref_c = admm_b200.admm_lasso(x, y).penalty(lam_c).parallel(world).opts(maxit=4000).fit()

    # a problem wide enough for the pipelined host ingest (p >= 1024), forced into several column panels
    os.environ["B200ADMM_PANEL_COLS"] = "256"
    n2, p2 = 4003, 1100
    x2 = np.asfortranarray(rng.normal(0.2, 2.0, size=(n2, p2)).astype(np.float32))
    b2 = np.zeros(p2); b2[:12] = rng.uniform(0.5, 1.5, size=12)
    y2 = (0.5 + x2 @ b2 + rng.normal(size=n2)).astype(np.float32)
    q0, qn = D.row_block(n2, world, rank)
    x2s, y2s = np.asfortranarray(x2[q0:q0 + qn]), y2[q0:q0 + qn].copy()
    ref2 = admm_b200.admm_lasso(x2, y2).penalty(nlambda=8).fit()

    # consensus on device-resident float32 row blocks with p >= 256: the fused DataStd + A'b + fp16-operand pass
    # (no standardised copy) that BASELINE config 5 runs at n = 1e6 x p = 8e4
    n3, p3 = 900 * world + 37, 320
    x3 = np.asfortranarray(rng.normal(0.1, 2.0, size=(n3, p3)).astype(np.float32))
    b3 = np.zeros(p3); b3[:9] = rng.uniform(0.5, 1.5, size=9)
    y3 = (0.3 + x3 @ b3 + rng.normal(size=n3)).astype(np.float32)
    lam3 = [0.2, 0.05]
    ref3 = admm_b200.admm_lasso(x3, y3).penalty(lam3).parallel(world).opts(maxit=4000).fit()
    s0, sn = D.row_block(n3, world, rank)
    X3d = torch.from_numpy(np.ascontiguousarray(x3[s0:s0 + sn].T)).cuda()      # (p, rows): the (rows, p) column-major block
    y3d = torch.from_numpy(y3[s0:s0 + sn].copy()).cuda()

    D.init_comm()
    f3 = admm_b200.admm_lasso(X3d.t(), y3d).penalty(lam3).parallel(world).opts(maxit=4000).fit()
    b3g, b3r = np.asarray(f3.beta.todense()), np.asarray(ref3.beta.todense())
    assert np.abs(b3g - b3r).max() < 2e-4, np.abs(b3g - b3r).max()
    assert np.abs(f3.niter.astype(int) - ref3.niter.astype(int)).max() <= max(3, 0.05 * ref3.niter.max()), (f3.niter, ref3.niter)
    t3 = torch.from_numpy(b3g.copy()).cuda()
    t30 = t3.clone(); dist.broadcast(t30, 0)
    assert torch.equal(t3, t30)                                           # identical on every rank

    f2 = admm_b200.admm_lasso(x2s, y2s).penalty(nlambda=8).fit()          # pipelined copy + per-panel all-reduces + sharded iterations
    b2g, b2r = np.asarray(f2.beta.todense()), np.asarray(ref2.beta.todense())
    tol2 = max(2e-4, 2e-5 * np.sqrt(p2))
    assert np.abs(b2g - b2r).max() < tol2, np.abs(b2g - b2r).max()
    assert abs(int(f2.niter.sum()) - int(ref2.niter.sum())) <= max(3, 0.03 * int(ref2.niter.sum())), (f2.niter, ref2.niter)
    del os.environ["B200ADMM_PANEL_COLS"]

    f = admm_b200.admm_lasso(xs, ys).penalty(nlambda=15).fit()
    bg, br = np.asarray(f.beta.todense()), np.asarray(ref.beta.todense())
    assert np.allclose(f.lambda_, ref.lambda_, rtol=1e-6), (f.lambda_[:3], ref.lambda_[:3])
    assert np.abs(bg - br).max() < 2e-4, np.abs(bg - br).max()
    assert abs(int(f.niter.sum()) - int(ref.niter.sum())) <= max(3, 0.03 * int(ref.niter.sum())), (f.niter, ref.niter)
    # sharded iterations: every rank returns the identical result ...
    t = torch.from_numpy(bg.copy()).cuda()
    t0 = t.clone(); dist.broadcast(t0, 0)
    assert torch.equal(t, t0)
    # ... and the replicated iterations (B200ADMM_SHARD_ITER=0, one-triangle kernel on every rank) agree with them
    os.environ["B200ADMM_SHARD_ITER"] = "0"
    fr = admm_b200.admm_lasso(xs, ys).penalty(nlambda=15).fit()
    del os.environ["B200ADMM_SHARD_ITER"]
    brr = np.asarray(fr.beta.todense())
    assert np.abs(brr - bg).max() < 1e-4, np.abs(brr - bg).max()
    assert np.abs(fr.niter.astype(int) - f.niter.astype(int)).max() <= 2, (fr.niter, f.niter)
    tr_ = torch.from_numpy(brr.copy()).cuda()
    tr0 = tr_.clone(); dist.broadcast(tr0, 0)
    assert torch.equal(tr_, tr0)

    fc = admm_b200.admm_lasso(xs, ys).penalty(lam_c).parallel(world).opts(maxit=4000).fit()
    bc, brc = np.asarray(fc.beta.todense()), np.asarray(ref_c.beta.todense())
    assert np.abs(bc - brc).max() < 2e-4, np.abs(bc - brc).max()
    assert np.abs(fc.niter.astype(int) - ref_c.niter.astype(int)).max() <= max(3, 0.05 * ref_c.niter.max()), (fc.niter, ref_c.niter)
    if rank == 0:
        from oracle import pyoracle as O
        o = O.lasso_path(x, y, lam_c, nthread=world, maxit=4000)
        assert np.abs(bc - o["beta"]).max() < 2e-4, np.abs(bc - o["beta"]).max()
        assert np.abs(fc.niter.astype(int) - o["niter"].astype(int)).max() <= max(3, 0.05 * o["niter"].max())
        print("mgpu_check ok: world=%d sharded-tall max|dbeta|=%.2e niter %d/%d; consensus max|dbeta vs oracle|=%.2e niter %s/%s"
              % (world, np.abs(bg - br).max(), int(f.niter.sum()), int(ref.niter.sum()),
                 np.abs(bc - o["beta"]).max(), fc.niter.tolist(), o["niter"].tolist()), flush=True)
    D.destroy_comm()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
