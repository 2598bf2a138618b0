"""Debug: accuracy of a <- a^-1 (float32) for the factorisation variants.  python tools/debug_spd_accuracy.py"""
import ctypes as C
import os
import subprocess
import sys
import numpy as np
sys.path.insert(0, ".")

if len(sys.argv) > 1:
    import torch
    from admm_b200 import _capi as K
    L = K.lib()
    for p in (1000, 1300, 4352, 10000):
        g = torch.Generator(device="cuda").manual_seed(p)
        n = 16 * p
        a = torch.zeros((p, p), device="cuda", dtype=torch.float64)
        for c0 in range(0, n, 16384):
            x = torch.randn((min(16384, n - c0), p), device="cuda", generator=g).double()
            a += x.t() @ x
        a += 0.3 * n * torch.eye(p, device="cuda", dtype=torch.float64)
        a32 = a.float().contiguous()
        a64 = a32.double()
        inv = a32.clone()
        work = torch.empty_like(inv)
        info = C.c_int(-1)
        torch.cuda.synchronize()
        K.check(L.b200admm_k_spd_inverse_f32(inv.data_ptr(), p, work.data_ptr(), C.byref(info)))
        r = inv.double() @ a64 - torch.eye(p, device="cuda", dtype=torch.float64)
        ref = torch.linalg.inv(a64)
        rel = ((inv.double() - ref).abs().max() / ref.abs().max()).item()
        rb = r.abs().amax(dim=1)
        print("%s p=%5d info=%d max|inv a - I| = %.3e  max rel err of inv = %.3e  worst rows %s sym %s"
              % (sys.argv[1], p, info.value, r.abs().max().item(), rel, torch.topk(rb, 3).indices.tolist(), bool(torch.equal(inv, inv.t()))), flush=True)
else:
    for mode in ("fast", "legacy"):
        env = dict(os.environ, B200ADMM_DIAG=mode)
        subprocess.run([sys.executable, sys.argv[0], mode], env=env)
    env = dict(os.environ, B200ADMM_FACTOR="legacy")
    subprocess.run([sys.executable, sys.argv[0], "cuda-core-path"], env=env)
