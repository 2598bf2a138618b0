"""Debug: where do the GPU and oracle LAD / BP traces part on bench.py's generator?
python tests/tools/debug_lad_trace.py [lad|bp] n p [maxit]"""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import torch
import admm_b200
from admm_b200 import _capi as K
from oracle import pyoracle as O

which = sys.argv[1] if len(sys.argv) > 1 else "lad"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
p = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
maxit = int(sys.argv[4]) if len(sys.argv) > 4 else 20
O.use_openblas(16)
O.omp_threads(16)
X = torch.empty((p, n), dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
if which == "lad":
    K.check(K.lib().b200admm_synth_f32(X.data_ptr(), y.data_ptr(), n, p, 0, 123, 0.0, 2.0, min(100, p), 1.0))
    Xd, yd = X.double(), y.double()
else:
    K.check(K.lib().b200admm_synth_f32(X.data_ptr(), y.data_ptr(), n, p, 0, 123, 0.0, 1.0, 0, 0.0))
    Xd = X.double()
    g = torch.Generator(device="cuda").manual_seed(123)
    bt = torch.zeros(p, dtype=torch.float64, device="cuda")
    idx = torch.randperm(p, device="cuda", generator=g)[:500]
    bt[idx] = torch.rand(500, dtype=torch.float64, device="cuda", generator=g)
    yd = Xd.t() @ bt
del X
with K.trace(which=0, cap=maxit + 5) as tr:
    m = admm_b200.admm_lad(Xd.t(), yd) if which == "lad" else admm_b200.admm_bp(Xd.t(), yd)
    f = m.opts(maxit=maxit).fit()
xh = Xd.cpu().numpy().T
yh = yd.cpu().numpy()
t0 = time.perf_counter()
o = O.lad(xh, yh, maxit=maxit, trace_cap=maxit + 5) if which == "lad" else O.bp(xh, yh, maxit=maxit, trace_cap=maxit + 5)
print(which, "n", n, "p", p, "oracle s %.1f" % (time.perf_counter() - t0), "niter", f.niter, o["niter"], flush=True)
tg, tc = tr.rows, o["trace"]
mm = min(len(tg), len(tc), maxit)
for r in range(mm):
    d = np.abs(tg[r] - tc[r]) / np.maximum(np.abs(tc[r]), 1e-300)
    d[(tg[r] == 0) & (tc[r] == 0)] = 0
    print("it %3d rel %.2e gpu %s" % (r, d.max(), np.array2string(tg[r], precision=14)))
    print("                    cpu %s" % (np.array2string(tc[r], precision=14)), flush=True)
if which == "lad":
    print("max|dbeta|", float(np.abs(f.beta - o["beta"]).max()))
