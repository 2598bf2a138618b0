"""Independent cross-checks of the CPU oracle (test infrastructure) against solvers that share no code with it:
scikit-learn's coordinate descent for the lasso / elastic net (the reference's README compares against glmnet,
which solves the same problem) and scipy's LP solver for LAD and basis pursuit.  The bands are the ones the
reference's README reports for ADMM against glmnet / quantreg (README.md:239-241, 287-289, 332, 363, 392).
The wide (n < p) solver has no reference-printed vector; this is its pin."""
import numpy as np
import pytest

from oracle import pyoracle as O


def readme_like(n, p, m, seed, mean=1.2):
    rng = np.random.default_rng(seed)
    b = np.zeros(p)
    b[:m] = rng.uniform(size=m)
    x = rng.normal(mean, 2.0, size=(n, p))
    y = 5.0 + x @ b + rng.normal(size=n)
    return np.asfortranarray(x), y


def sklearn_glmnet_equivalent(x, y, lam, alpha=1.0):
    """glmnet(standardize = TRUE) through scikit-learn: columns scaled by their population sd, the penalty
    lam * (alpha |b|_1 + (1 - alpha)/2 |b|_2^2) applied to the standardised coefficients (y on its own scale)."""
    from sklearn.linear_model import ElasticNet, Lasso
    mu, sd = x.mean(axis=0), x.std(axis=0)
    xs = (x - mu) / sd
    if alpha == 1.0:
        m = Lasso(alpha=lam, fit_intercept=True, tol=1e-12, max_iter=200000)
    else:
        m = ElasticNet(alpha=lam, l1_ratio=alpha, fit_intercept=True, tol=1e-12, max_iter=200000)
    m.fit(xs, y)
    coef = m.coef_ / sd
    return np.concatenate([[m.intercept_ - coef @ mu], coef])


@pytest.mark.parametrize("n,p,lam", [(300, 40, 0.1), (300, 40, 0.02)])
def test_tall_lasso_against_coordinate_descent(n, p, lam):
    x, y = readme_like(n, p, 6, seed=1)
    o = O.lasso_path(x, y, [lam])
    ref = sklearn_glmnet_equivalent(x, y, lam)
    assert np.abs(o["beta"][:, 0] - ref).max() < 6e-4                  # README n > p band: [-2.9e-4, 7.3e-5] at n = 1e4
    assert np.array_equal(o["beta"][1:, 0] != 0, np.abs(ref[1:]) > 1e-10) or np.abs(o["beta"][1:, 0] - ref[1:])[(o["beta"][1:, 0] != 0) != (np.abs(ref[1:]) > 1e-10)].max() < 6e-4


@pytest.mark.parametrize("alpha", [0.5, 0.3])
def test_tall_enet_against_coordinate_descent(alpha):
    """glmnet (gaussian family) -- and the reference, which standardises y as well (src/DataStd.h:103-127) --
    apply the elastic-net penalty on the scale where y has unit variance; the l1 part is scale-free, the ridge
    part is not.  scikit-learn on (x_std, y / sd_y) with alpha = lam / sd_y solves that problem."""
    from sklearn.linear_model import ElasticNet
    x, y = readme_like(300, 40, 6, seed=2)
    lam = 0.1
    o = O.lasso_path(x, y, [lam], model="enet", alpha=alpha)
    mu, sd, my, sy = x.mean(axis=0), x.std(axis=0), y.mean(), y.std()
    m = ElasticNet(alpha=lam / sy, l1_ratio=alpha, fit_intercept=False, tol=1e-12, max_iter=200000)
    m.fit((x - mu) / sd, (y - my) / sy)
    coef = m.coef_ * sy / sd
    ref = np.concatenate([[my - coef @ mu], coef])
    assert np.abs(o["beta"][:, 0] - ref).max() < 6e-4                  # README n > p band for enet: [-2.2e-4, 8.2e-5]


@pytest.mark.parametrize("n,p,lam", [(60, 200, 0.3), (80, 300, 0.15)])
def test_wide_lasso_against_coordinate_descent(n, p, lam):
    x, y = readme_like(n, p, 8, seed=3, mean=0.0)
    o = O.lasso_path(x, y, [lam], maxit=20000)
    ref = sklearn_glmnet_equivalent(x, y, lam)
    assert int(o["niter"][0]) <= 20000
    assert np.abs(o["beta"][:, 0] - ref).max() < 6e-3                  # README p > n band: [-1.5e-3, 2.1e-3] at n = 1e3
    big = np.abs(ref[1:]) > 2e-2
    assert (o["beta"][1:, 0][big] != 0).all()                          # every clearly active variable is found


def test_lad_against_linear_programming():
    from scipy.optimize import linprog
    rng = np.random.default_rng(4)
    n, p = 120, 8
    x = np.asfortranarray(rng.normal(size=(n, p)))
    y = x @ rng.uniform(size=p) + rng.standard_t(df=2, size=n)
    o = O.lad(x, y, intercept=False)
    # min sum(u + v)  s.t.  x b + u - v = y, u, v >= 0
    c = np.concatenate([np.zeros(p), np.ones(2 * n)])
    A = np.hstack([x, np.eye(n), -np.eye(n)])
    res = linprog(c, A_eq=A, b_eq=y, bounds=[(None, None)] * p + [(0, None)] * (2 * n), method="highs")
    assert res.status == 0
    assert np.abs(o["beta"][1:] - res.x[:p]).max() < 2e-2               # README band vs quantreg: +-7e-3 (n = 1e3)
    # and the objective is as good as the LP's to the solver's tolerance
    obj = np.abs(y - x @ o["beta"][1:]).sum()
    assert obj <= res.fun * (1 + 1e-3)


def test_bp_against_linear_programming():
    from scipy.optimize import linprog
    rng = np.random.default_rng(5)
    n, p, k = 40, 100, 6
    A = np.asfortranarray(rng.normal(size=(n, p)))
    bt = np.zeros(p)
    bt[rng.choice(p, k, replace=False)] = rng.uniform(0.5, 1.5, size=k)
    b = A @ bt
    o = O.bp(A, b)
    c = np.ones(2 * p)
    res = linprog(c, A_eq=np.hstack([A, -A]), b_eq=b, bounds=[(0, None)] * (2 * p), method="highs")
    assert res.status == 0
    x_lp = res.x[:p] - res.x[p:]
    assert np.abs(x_lp - bt).max() < 1e-8                               # exact recovery of the sparse signal by the LP
    assert np.abs(np.asarray(o["beta"]).ravel() - bt).max() < 5e-3      # README recovery band: [-1.3e-3, 2.1e-3]
