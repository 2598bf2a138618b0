"""Launches tests/mgpu_check.py with one rank per visible GPU (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_sharded_paths_on_all_visible_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % n,
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mgpu_check ok" in r.stdout
