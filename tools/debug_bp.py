import sys
import numpy as np
sys.path.insert(0, ".")
import admm_b200 as A
from oracle import pyoracle as O
rng = np.random.default_rng(5)
n, p, k = 120, 500, 12
x = np.asfortranarray(rng.normal(size=(n, p)))
bt = np.zeros(p)
bt[rng.choice(p, k, replace=False)] = rng.uniform(0.5, 1.5, size=k)
y = x @ bt
for eps in (1e-4, 1e-6):
    f = A.admm_bp(x, y).opts(eps_abs=eps, eps_rel=eps).fit()
    o = O.bp(x, y, eps_abs=eps, eps_rel=eps)
    b = np.asarray(f.beta.todense())[:, 0]
    print("eps", eps, "niter gpu/cpu", f.niter, o["niter"], "max|b-bo|", np.abs(b - o["beta"]).max(),
          "feas gpu/cpu", np.abs(x @ b - y).max(), np.abs(x @ o["beta"] - y).max(),
          "rec gpu/cpu", np.abs(b - bt).max(), np.abs(o["beta"] - bt).max(), "rho", f.info["rho"])
from admm_b200 import _capi as K
with K.trace(which=0, cap=20000) as tr:
    f = A.admm_bp(x, y).opts(eps_abs=1e-6, eps_rel=1e-6).fit()
o = O.bp(x, y, eps_abs=1e-6, eps_rel=1e-6, trace_cap=20000)
tg, tc = tr.rows, o["trace"][:o["niter"]]
m = min(len(tg), len(tc))
rel = np.abs(tg[:m] / np.where(tc[:m] == 0, 1, tc[:m]) - 1).max(axis=1)
bad = np.argmax(rel > 1e-6) if (rel > 1e-6).any() else -1
print("first iteration with trace rel diff > 1e-6:", bad, "of", m)
if bad >= 0:
    print(tg[max(0, bad - 1):bad + 2]); print(tc[max(0, bad - 1):bad + 2])
