#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for NS in 1 2 4 8; do
B200ADMM_GEMV_NSEG=$NS timeout 300 python - <<'P'
import sys, os, ctypes as C
sys.path.insert(0, ".")
import torch
from admm_b200 import _capi as K
L = K.lib()
st = torch.cuda.ExternalStream(L.b200admm_stream())
for (m, nc) in ((80000, 80000), (40000, 40000)):
    a = torch.randn((nc, m), device="cuda", dtype=torch.float32)
    v = torch.randn(m, device="cuda", dtype=torch.float32)
    out = torch.empty(nc, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        K.check(L.b200admm_k_gemv_t_f32(a.data_ptr(), m, nc, v.data_ptr(), out.data_ptr()))
        e1.record(st); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ref = (a[:64].double() @ v.double()).float()
    print("nseg %s gemv_t f32 m=%d ncol=%d: best %.3f ms %.1f GB/s  err %.2e" % (os.environ["B200ADMM_GEMV_NSEG"], m, nc, min(ts), 4.0 * m * nc / min(ts) / 1e6, float((out[:64] - ref).abs().max())), flush=True)
    del a, v, out
P
done
