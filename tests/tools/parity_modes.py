"""GPU fit vs CPU oracle for the tall test problems under each Gram kernel (run on the GPU box, one
process per mode because the mode is read from the environment):
    B200ADMM_GRAM=tf32 python tests/tools/parity_modes.py ; python tests/tools/parity_modes.py"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import admm_b200
from oracle import pyoracle as O


def make_problem(n, p, seed, nsig=10, mean=0.0):
    rng = np.random.default_rng(seed)
    x = rng.normal(mean, 2.0, size=(n, p))
    b = np.zeros(p)
    b[:nsig] = rng.uniform(size=nsig)
    y = x @ b + rng.normal(size=n)
    return np.asfortranarray(x), y


for n, p, model, alpha in [(2000, 200, "lasso", 1.0), (3000, 256, "enet", 0.5), (3000, 256, "lasso", 1.0), (6000, 512, "lasso", 1.0),
                           (6000, 512, "enet", 0.5), (20000, 1024, "lasso", 1.0)]:
    x, y = make_problem(n, p, seed=n + p)
    if model == "lasso":
        f = admm_b200.admm_lasso(x, y).penalty(nlambda=25).fit()
    else:
        f = admm_b200.admm_enet(x, y).penalty(nlambda=25, alpha=alpha).fit()
    o = O.lasso_path(x, y, nlambda=25, model=model, alpha=alpha)
    bg, bc = np.asarray(f.beta.todense()), o["beta"]
    d = np.abs(bg - bc).max(axis=0)
    print(os.environ.get("B200ADMM_GRAM", "default"), n, p, model, "max|dbeta| %.2e at lambda %d; per-lambda x1e4:" % (d.max(), d.argmax()),
          np.round(d * 1e4, 2).tolist(), "niter", int(f.niter.sum()), int(o["niter"].sum()), flush=True)
