"""Time a <- a^-1 (float32, p x p) in isolation for the factorisation variants: python tools/time_factor.py [p]"""
import ctypes as C
import os
import subprocess
import sys
sys.path.insert(0, ".")

if len(sys.argv) > 2:
    import torch
    from admm_b200 import _capi as K
    L = K.lib()
    p = int(sys.argv[1])
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((4 * p, p), device="cuda", generator=g)
    a0 = (x.t() @ x + 0.5 * 4 * p * torch.eye(p, device="cuda")).contiguous()
    del x
    a = a0.clone()
    work = torch.empty_like(a)
    info = C.c_int(0)
    st = torch.cuda.ExternalStream(L.b200admm_stream())
    times = []
    for rep in range(6):
        a.copy_(a0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        K.check(L.b200admm_k_spd_inverse_f32(a.data_ptr(), p, work.data_ptr(), C.byref(info)))
        e1.record(st)
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    print("%-10s p=%d  ms per call: %s" % (sys.argv[2], p, " ".join("%.2f" % t for t in times)), flush=True)
else:
    p = sys.argv[1] if len(sys.argv) > 1 else "10000"
    for mode, env in (("fast", {"B200ADMM_DIAG": "fast"}), ("legacydiag", {"B200ADMM_DIAG": "legacy"}), ("cudacore", {"B200ADMM_FACTOR": "legacy"})):
        subprocess.run([sys.executable, sys.argv[0], p, mode], env=dict(os.environ, **env))
