"""python tools/time_diag.py -- ms per launch of the one-CTA diagonal-block kernels (0 empty, 1 old, 2 fast)"""
import ctypes as C, sys
sys.path.insert(0, ".")
import torch
from admm_b200 import _capi as K
L = K.lib()
L.b200admm_debug_diag_ms.restype = C.c_double
L.b200admm_debug_diag_ms.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
for lda in (128, 10000):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((512, 128), device="cuda", generator=g)
    a = torch.zeros((128, lda), device="cuda")
    a[:, :128] = x.t() @ x + 64 * torch.eye(128, device="cuda")
    dinv = torch.zeros((128, 128), device="cuda")
    for mode in (0, 1, 2, 0, 1, 2):
        b = a.clone()
        torch.cuda.synchronize()
        ms = L.b200admm_debug_diag_ms(mode, 1, b.data_ptr(), lda, dinv.data_ptr())     # one launch on fresh data
        b = a.clone()
        torch.cuda.synchronize()
        ms50 = L.b200admm_debug_diag_ms(mode, 50, b.data_ptr(), lda, dinv.data_ptr())  # 50 back to back (refactors its own output: timing only)
        print("lda %5d mode %d: single launch %.1f us, mean of 50 %.1f us" % (lda, mode, ms * 1e3, ms50 * 1e3), flush=True)
