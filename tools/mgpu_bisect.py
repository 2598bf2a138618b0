"""Debug (multi-GPU): which switch makes the small-n pipelined sharded fit go wrong on 8 ranks?
torchrun --nproc-per-node N tools/mgpu_bisect.py"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


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
    cases = []
    for (n2, p2) in ((4003, 1100), (1000 * world + 3, 1100), (4003, 1024), (16003, 1100)):
        x2 = np.asfortranarray(rng.normal(0.2, 2.0, size=(n2, p2)).astype(np.float32))
        b2 = np.zeros(p2); b2[:12] = rng.uniform(0.5, 1.5, size=12)
        y2 = (0.5 + x2 @ b2 + rng.normal(size=n2)).astype(np.float32)
        ref2 = admm_b200.admm_lasso(x2, y2).penalty(nlambda=8).fit()
        cases.append((n2, p2, x2, y2, ref2))
    D.init_comm()
    for (n2, p2, x2, y2, ref2) in cases:
        q0, qn = D.row_block(n2, world, rank)
        x2s, y2s = np.asfortranarray(x2[q0:q0 + qn]), y2[q0:q0 + qn].copy()
        b2r = np.asarray(ref2.beta.todense())
        for env in ({"B200ADMM_PANEL_COLS": "256"}, {"B200ADMM_PANEL_COLS": "256", "B200ADMM_SHARD_ITER": "0"}, {"B200ADMM_PIPELINE": "0"},
                    {"B200ADMM_PIPELINE": "0", "B200ADMM_SHARD_ITER": "0"}, {}, {"B200ADMM_GRAM": "tf32"}):
            for k, v in env.items():
                os.environ[k] = v
            try:
                f2 = admm_b200.admm_lasso(x2s, y2s).penalty(nlambda=8).fit()
                b2g = np.asarray(f2.beta.todense())
                msg = "max|dbeta| %.3e niter %s vs %s rho %.6g vs %.6g" % (np.abs(b2g - b2r).max(), f2.niter.tolist(), ref2.niter.tolist(), f2.info["rho"], ref2.info["rho"])
            except Exception as ex:
                msg = "ERROR " + repr(ex)[:200]
            for k in env:
                del os.environ[k]
            if rank == 0:
                print("n %d p %d rows/rank %d %-70s %s" % (n2, p2, qn, str(env), msg), flush=True)
    D.destroy_comm()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
