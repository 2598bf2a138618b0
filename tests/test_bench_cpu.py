"""bench.py's reference arm measures the CPU iteration phase on a Gram matrix with the full-size design's
statistics (wishart_problem) instead of forming the 40 GB design.  Check on a size the oracle can do both ways
that the surrogate reproduces the real design's path: same lambda_max scale, iteration count within a few per cent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_wishart_surrogate_reproduces_the_iteration_count():
    import bench
    from oracle import pyoracle as O
    n, p, nl = 40000, 400, 30
    rng = np.random.default_rng(3)
    x = np.asfortranarray((rng.standard_normal((n, p)) * 2.0).astype(np.float32))
    beta = np.zeros(p, dtype=np.float32)
    beta[:100] = rng.uniform(size=100)
    y = (x @ beta + rng.standard_normal(n).astype(np.float32)).astype(np.float32)
    O.standardize_f32(x, y)
    G = O.gram_tn_f32(x)
    xy = (x.T.astype(np.float64) @ y.astype(np.float64)).astype(np.float32)
    lam0 = float(np.abs(xy).max())
    grid = np.exp(np.linspace(np.log(lam0), np.log(lam0 * 1e-4), nl))
    real = O.tall_path_from_gram(G, xy, grid)

    Gs, xys = bench.wishart_problem(n, p, seed=7)
    lam0s = float(np.abs(xys).max())
    grids = np.exp(np.linspace(np.log(lam0s), np.log(lam0s * 1e-4), nl))
    sur = O.tall_path_from_gram(Gs, xys, grids)

    assert abs(lam0s / lam0 - 1.0) < 0.1                                  # same lambda_max up to sampling noise
    nr, ns = int(real["niter"].sum()), int(sur["niter"].sum())
    assert abs(ns - nr) <= 0.08 * nr, (nr, ns)
    assert abs(float(sur["rho"]) / float(real["rho"]) - 1.0) < 0.1
