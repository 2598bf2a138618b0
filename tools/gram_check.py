"""Accuracy + speed of the Gram kernels on a synthetic design (run on the GPU box).
usage: python tools/gram_check.py [n] [p]"""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from admm_b200 import _capi as K

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
L = K.lib()
X = torch.empty((p, n), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
K.check(L.b200admm_synth_f32(X.data_ptr(), None, n, p, 0, 7, 0.3, 2.0, 10, 1.0))
ref = None
if n * p <= 2_000_000_000:
    X64 = X.double()
    ref = X64 @ X64.t()
    del X64
    dscale = torch.sqrt(torch.outer(torch.diag(ref), torch.diag(ref)))
MODES = [int(m) for m in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3]
for mode, name in ((0, "cuda-core fp32"), (1, "tcgen05 3xTF32 (trunc split)"), (2, "tcgen05 3xTF32 (rn split)"), (3, "tcgen05 3xFP16 (rn split)")):
    if mode not in MODES:
        continue
    G = torch.zeros((p, p), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    K.check(L.b200admm_k_gram_f32(X.data_ptr(), n, p, G.data_ptr(), mode))     # warm-up
    t0 = time.perf_counter()
    K.check(L.b200admm_k_gram_f32(X.data_ptr(), n, p, G.data_ptr(), mode))
    dt = time.perf_counter() - t0
    msg = "%-32s %8.3f ms  %7.1f TFLOP/s (syrk flops n p (p+1))" % (name, dt * 1e3, n * p * (p + 1) / dt / 1e12)
    if ref is not None:
        err = (G.double() - ref).abs() / dscale
        dg = torch.diag(G).double() / torch.diag(ref) - 1
        msg += "  max rel err %.2e  rms %.2e  diag bias mean %.2e max|.| %.2e sym %s" % (
            err.max().item(), err.pow(2).mean().sqrt().item(), dg.mean().item(), dg.abs().max().item(),
            bool(torch.equal(G, G.t())))
    print(msg, flush=True)
