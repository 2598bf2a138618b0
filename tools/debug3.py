import sys
import numpy as np
sys.path.insert(0, ".")
import admm_b200 as A
from admm_b200 import _capi as K
from oracle import pyoracle as O
np.set_printoptions(linewidth=220, precision=17)
for seed in (5, 6, 7):
    rng = np.random.default_rng(seed)
    n, p, k = 120, 500, 12
    x = np.asfortranarray(rng.normal(size=(n, p)))
    bt = np.zeros(p); bt[rng.choice(p, k, replace=False)] = rng.uniform(0.5, 1.5, size=k)
    y = x @ bt
    with K.trace(which=0, cap=2000) as tr:
        f = A.admm_bp(x, y).fit()
    o = O.bp(x, y, trace_cap=2000)
    m = min(f.niter, o["niter"])
    d = np.abs(tr.rows[:m] - o["trace"][:m]) / np.maximum(np.abs(o["trace"][:m]), 1e-300)
    bad = int(np.argmax(d.max(axis=1) > 1e-9)) if (d.max(axis=1) > 1e-9).any() else -1
    print("seed", seed, "niter", f.niter, o["niter"], "first row with rel diff > 1e-9:", bad)
    if bad >= 0:
        for r in range(max(0, bad - 2), min(m, bad + 2)):
            print(r, "gpu", tr.rows[r]); print(r, "cpu", o["trace"][r])
