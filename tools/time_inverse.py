import ctypes as C, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from admm_b200 import _capi as K
L = K.lib()
for p in [int(a) for a in sys.argv[1:]] or [4992, 5000, 5120, 10000]:
    g = torch.Generator(device="cuda").manual_seed(p)
    x = torch.randn((p, 2 * p), device="cuda", generator=g)
    a0 = x @ x.t() + 5.0 * torch.eye(p, device="cuda")
    work = torch.empty((p, p), device="cuda")
    info = C.c_int(0)
    ts = []
    for rep in range(3):
        a = a0.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K.check(L.b200admm_k_spd_inverse_f32(a.data_ptr(), p, work.data_ptr(), C.byref(info)))
        ts.append(time.perf_counter() - t0)
    err = float(((a.double() @ a0.double()) - torch.eye(p, device="cuda", dtype=torch.float64)).abs().max())
    print("p=%d  spd_inverse times %s  max|K^-1 K - I|=%.2e" % (p, ["%.1f ms" % (t * 1e3) for t in ts], err), flush=True)
