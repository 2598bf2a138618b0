"""Time the float64 products of the LAD / BP setup (tensor-core vs CUDA-core tiles): python tools/time_gemm_f64.py"""
import ctypes as C
import os
import subprocess
import sys
sys.path.insert(0, ".")

if len(sys.argv) > 1:
    import torch
    from admm_b200 import _capi as K
    L = K.lib()
    for (name, ta, tb, M, N, Kd, mode) in (("LAD Gram X'X  n=5e5 p=5e3", 1, 0, 5000, 5000, 500000, 3),
                                           ("BP Gram AA'   n=5e3 p=5e5", 0, 1, 5000, 5000, 500000, 3),
                                           ("BP M = L^-1 A (one 5000 x 5000 x 100000 panel product)", 0, 0, 5000, 100000, 5000, 4)):
        a = torch.randn((M if ta else Kd, Kd if ta else M), device="cuda", dtype=torch.float64)      # column-major (rows x cols).T
        b = torch.randn((Kd if tb else N, N if tb else Kd), device="cuda", dtype=torch.float64)
        c = torch.zeros((N, M), device="cuda", dtype=torch.float64)
        ms = C.c_float(0)
        lda = Kd if ta else M
        ldb = N if tb else Kd
        K.check(L.b200admm_k_gemm_f64(ta, tb, M, N, Kd, 1.0, a.data_ptr(), lda, b.data_ptr(), ldb, 0.0, c.data_ptr(), M, mode, C.byref(ms), 1))
        K.check(L.b200admm_k_gemm_f64(ta, tb, M, N, Kd, 1.0, a.data_ptr(), lda, b.data_ptr(), ldb, 0.0, c.data_ptr(), M, mode, C.byref(ms), 2))
        flop = 2.0 * M * N * Kd * (0.5 if mode & 1 else 1.0) * (0.5 if mode & 4 else 1.0)
        print("%-10s %-58s %8.1f ms  %6.1f TFLOP/s" % (sys.argv[1], name, ms.value, flop / ms.value / 1e9), flush=True)
        del a, b, c
else:
    for mode, env in (("tensor", {}), ("simt", {"B200ADMM_GEMM_F64": "simt"})):
        subprocess.run([sys.executable, sys.argv[0], mode], env=dict(os.environ, **env))
