import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "tests/golden")
import admm_b200 as A
from admm_b200 import _capi as K
from oracle import pyoracle as O
np.set_printoptions(linewidth=200, precision=10)
rng = np.random.default_rng(5)
n, p, k = 120, 500, 12
x = np.asfortranarray(rng.normal(size=(n, p)))
bt = np.zeros(p); bt[rng.choice(p, k, replace=False)] = rng.uniform(0.5, 1.5, size=k)
y = x @ bt
with K.trace(which=0, cap=2000) as tr:
    f = A.admm_bp(x, y).fit()
o = O.bp(x, y, trace_cap=2000)
m = min(f.niter, o["niter"], 40)
d = np.abs(tr.rows[:m] - o["trace"][:m]) / np.maximum(np.abs(o["trace"][:m]), 1e-300)
print("BP niter", f.niter, o["niter"], "max rel diff per column", d.max(axis=0), "first bad row", np.argmax(d.max(axis=1) > 1e-7))
print(tr.rows[:3]); print(o["trace"][:3])
# consensus
def problem(n, p, seed, nsig=8, noise=1.0):
    rng = np.random.default_rng(seed)
    x = rng.normal(0.0, 2.0, size=(n, p)); b = np.zeros(p); b[:nsig] = rng.uniform(0.5, 1.5, size=nsig)
    return np.asfortranarray(x), x @ b + noise * rng.normal(size=n), b
x, y, _ = problem(400, 30, seed=3)
with K.trace(which=0, cap=500) as tr:
    f = A.admm_lasso(x, y).penalty([0.2]).parallel(2).opts(maxit=40).fit()
o = O.lasso_path(x, y, [0.2], nthread=2, maxit=40, trace_lambda=0, trace_cap=500)
print("cons niter", f.niter, o["niter"], "rows", len(tr.rows))
d = np.abs(tr.rows[:40] - o["trace"][:40]) / np.maximum(np.abs(o["trace"][:40]), 1e-300)
print("max rel diff per column", d.max(axis=0))
print(tr.rows[:3]); print(o["trace"][:3]); print(tr.rows[37:40]); print(o["trace"][37:40])
print("beta diff", np.abs(np.asarray(f.beta.todense())[:, 0] - o["beta"][:, 0]).max())
