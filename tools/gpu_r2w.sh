#!/bin/bash
# r2w (1 GPU): GPU tests + smoke after the last wide-solver changes; gemv_t at the consensus shape (8e4 x 8e4)
set -u
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > $O/r2w_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2w_pytest.log
tail -n 3 $O/r2w_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2w_smoke.log 2>&1; tail -n 1 $O/r2w_smoke.log
timeout 300 python - > $O/r2w_gemv_consensus_shape.log 2>&1 <<'P'
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from admm_b200 import _capi as K
L = K.lib()
st = torch.cuda.ExternalStream(L.b200admm_stream())
for (m, nc) in ((80000, 80000), (20000, 20000), (500000, 5000)):
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
    print("gemv_t f32 m=%d ncol=%d: ms %s  best %.1f GB/s" % (m, nc, " ".join("%.3f" % t for t in ts), 4.0 * m * nc / min(ts) / 1e6), flush=True)
    del a, v, out
P
cat $O/r2w_gemv_consensus_shape.log
