"""world_size-2 gloo tests (CPU) of the host-side plumbing of the row-sharded paths: the reference's
row split and the byte broadcast used to hand the NCCL identifier to every rank."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_block_is_the_reference_split():
    from admm_b200.dist import row_block
    for n, N in [(100, 2), (101, 2), (1000, 3), (1_000_000, 8), (17, 5)]:
        chunk = n // N
        rows = [row_block(n, N, r) for r in range(N)]
        assert rows[0][0] == 0
        for r in range(N - 1):
            assert rows[r] == (r * chunk, chunk)                       # PADMMLasso.h:169-172
        assert rows[-1] == ((N - 1) * chunk, chunk + n % N)            # PADMMLasso.h:173-177
        assert sum(c for _, c in rows) == n
    with pytest.raises(ValueError):
        row_block(3, 5, 0)
    with pytest.raises(ValueError):
        row_block(10, 2, 2)


def test_gloo_world2_broadcast_and_partition(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent('''
        import os, sys
        sys.path.insert(0, %r)
        import numpy as np
        import torch.distributed as dist
        from admm_b200.dist import broadcast_bytes, row_block
        dist.init_process_group(backend="gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = bytes(range(128)) if rank == 0 else bytes(128)
        got = broadcast_bytes(ident, 0)
        assert got == bytes(range(128)), got[:8]
        # every rank takes its block of the same global problem; the union is the problem
        n = 1001
        r0, nr = row_block(n, world, rank)
        import torch
        cover = torch.zeros(n, dtype=torch.int32)
        cover[r0:r0 + nr] = 1
        dist.all_reduce(cover)
        assert int(cover.min()) == 1 and int(cover.max()) == 1
        # global column statistics from row shards = the single-process statistics (what DataStd needs)
        rng = np.random.default_rng(0)
        x = rng.normal(1.0, 2.0, size=(n, 7))
        s = torch.from_numpy(x[r0:r0 + nr].sum(axis=0)); dist.all_reduce(s)
        mean = s.numpy() / n
        ss = torch.from_numpy(((x[r0:r0 + nr] - mean) ** 2).sum(axis=0)); dist.all_reduce(ss)
        assert np.allclose(mean, x.mean(axis=0)) and np.allclose(np.sqrt(ss.numpy() / n), x.std(axis=0))
        dist.destroy_process_group()
        print("rank", rank, "ok")
    ''' % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
