"""Pin the CPU oracle against the only results the reference ever published: the coefficient
columns printed in /root/reference/README.md (copied verbatim into tests/golden/readme_vectors.py),
on inputs regenerated bit-for-bit from R's set.seed(123) stream (tests/golden/make_readme_data.py).

Tolerances: the README prints 9-10 significant digits.  f64 paths (LAD, BP) reproduce to ~1e-10.
f32 paths reproduce to ~1e-6 -- for the serial lasso / elastic-net columns once the ONE difference between the build that
knitted the README and the source as it stands is taken into account:

  The default rho of the tall solver is ev^(1/3) lambda^(2/3) with ev the UNCONVERGED Ritz value of
  Spectra::SymEigsSolver(op, nev = 1, ncv = 3).compute(10, 0.1) (src/ADMMLassoTall.h:194-202, inherited by ADMMEnetTall).
  With that call (ev = 178.950 on the README's X, 2.7 % below lambda_max = 183.839) the oracle reproduces the ENET column
  to 1.9e-6 but the LASSO column only to 1.2e-5 -- and not because of the stopping iteration (both runs stop at iteration
  31 with a 3 % margin).  Every tall-solver output the README holds is reproduced, at its print precision, by the SAME
  restatement with **ncv = 2** in that one call (ev = 181.626):
      README lasso column (README.md:66-88)            9.5e-7   (ncv = 3: 1.2e-5)
      README elastic-net column (README.md:101-122)    4.8e-7   (ncv = 3: 1.9e-6)
      README n = 1e4, p = 1e3 benchmark, minimum of range(coef(glmnet) - admm_lasso)   2.0e-7   (ncv = 3: 2.7e-5)
      the same for admm_enet(alpha = 0.6)                                              1.0e-7   (ncv = 3: 1.2e-4)
  (the last two in tests/test_oracle_readme_benchmarks.py; a rho scan puts the benchmark's two minima, independently, at
  0.9950 +- 0.0003 of the ncv = 3 value, ncv = 2 gives 0.9947; ncv = 3 with other tolerances, ncv = 4..6, the converged
  lambda_max and other restart rules fit at most one of the four).  So the README was knitted by a build whose Lanczos call
  kept TWO basis vectors -- an earlier revision of src/ADMMLassoTall.h:196 -- and the source as it stands keeps three.
  The oracle and the CUDA library follow today's source (ncv = 3); `pyoracle.lanczos_ncv(2)` exists for these forensics
  only.  1.2e-5 is therefore the distance between two correct runs of the reference 0.5 % apart in rho, not noise to be
  covered by a tolerance.  The p > n rows of the wide solver tell the same story for its gamma (README build: the ncv = 5
  estimate; tests/test_oracle_readme_benchmarks.py); parallel lasso, LAD and BP use no such estimate.

The iteration counts (31 / 339 / 22 / 443 / 72) are those of the survey's independent NumPy probes.
"""
import os

import numpy as np
import pytest

import readme_vectors as R
from oracle import pyoracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAM = float(np.exp(-2))


@pytest.fixture(scope="module")
def lasso_xy():
    d = np.load(os.path.join(G, "readme_lasso_data.npz"))
    return d["x"], d["y"]


def test_r_rng_known_answers():
    import make_readme_data as M
    r = M.RRng(123)
    assert np.allclose(r.runif(3), [0.2875775201246142, 0.7883051354438066, 0.4089769218116999], atol=1e-15)
    x, y, b = M.lasso_data()
    d = np.load(os.path.join(G, "readme_lasso_data.npz"))
    assert np.array_equal(x, d["x"]) and np.array_equal(y, d["y"])


def test_lasso_readme_column(lasso_xy):
    x, y = lasso_xy
    r = O.lasso_path(x, y, [LAM])
    b = r["beta"][:, 0]
    assert r["niter"][0] == 31
    assert np.array_equal(b != 0, R.LASSO_ADMM != 0)          # support bit-exact
    assert np.abs(b - R.LASSO_ADMM).max() < 2e-5
    assert np.abs(b - R.LASSO_GLMNET).max() < 3e-4            # README.md:239 band vs glmnet


def test_lasso_readme_column_exact_rho(lasso_xy):
    """The converged lambda_max happens to fit the lasso column too (its rho is 0.4 % above the ncv = 2 one) -- but it misses
    the elastic-net column by 5e-5 and the README's benchmark minima by 2.6e-4; ncv = 2 fits all four (module docstring)."""
    x, y = lasso_xy
    xs = np.asfortranarray(x, dtype=np.float32).copy(order="F")
    ys = y.astype(np.float32)
    st = O.standardize_f32(xs, ys)
    Gm = O.gram_tn_f32(xs).astype(np.float64)
    ev = np.linalg.eigvalsh(np.tril(Gm) + np.tril(Gm, -1).T).max()
    il = np.float32(LAM * 100 / st["scaleY"])
    rho = ev ** (1 / 3) * float(il) ** (2 / 3)
    r = O.lasso_path(x, y, [LAM], rho=rho)
    assert np.abs(r["beta"][:, 0] - R.LASSO_ADMM).max() < 2e-6


def test_readme_tall_columns_were_knitted_with_two_lanczos_vectors(lasso_xy):
    """See the module docstring: with ncv = 2 in the Spectra call both tall columns are reproduced at the README's print
    precision; with the source's ncv = 3 the lasso column is 1.2e-5 away although the stopping iteration is not at issue."""
    x, y = lasso_xy
    lasso3 = O.lasso_path(x, y, [LAM])
    enet3 = O.lasso_path(x, y, [LAM], model="enet", alpha=0.5)
    with O.lanczos_ncv(2):
        lasso2 = O.lasso_path(x, y, [LAM])
        enet2 = O.lasso_path(x, y, [LAM], model="enet", alpha=0.5)
    again = O.lasso_path(x, y, [LAM])                                      # the switch does not leak
    assert again["eig"] == lasso3["eig"] == enet3["eig"] and lasso2["eig"] == enet2["eig"]
    assert abs(lasso3["eig"] - 178.950) < 1e-3 and abs(lasso2["eig"] - 181.626) < 1e-3
    err = lambda r, ref: float(np.abs(r["beta"][:, 0] - ref).max())
    print("\n[readme] lasso column: ncv = 3 %.2e, ncv = 2 %.2e;  enet column: ncv = 3 %.2e, ncv = 2 %.2e"
          % (err(lasso3, R.LASSO_ADMM), err(lasso2, R.LASSO_ADMM), err(enet3, R.ENET_ADMM), err(enet2, R.ENET_ADMM)))
    assert err(lasso2, R.LASSO_ADMM) < 1.2e-6 and err(enet2, R.ENET_ADMM) < 1.2e-6       # both at print precision
    assert err(lasso3, R.LASSO_ADMM) > 8e-6 and err(enet3, R.ENET_ADMM) < 3e-6            # today's source: enet only
    assert lasso3["niter"][0] == lasso2["niter"][0] == 31 and enet3["niter"][0] == enet2["niter"][0] == 22
    for r, ref in ((lasso2, R.LASSO_ADMM), (enet2, R.ENET_ADMM)):
        assert np.array_equal(r["beta"][:, 0] != 0, ref != 0)
    # not a stopping-rule accident: the last iteration passes with margin, the one before fails clearly
    o = O.lasso_path(x, y, [LAM], trace_lambda=0, trace_cap=64)
    last, prev = o["trace"][30], o["trace"][29]
    assert last[1] < 0.8 * last[0] and last[3] < 0.98 * last[2] and prev[3] > 1.3 * prev[2]


def test_parallel_lasso_readme_column(lasso_xy):
    x, y = lasso_xy
    r = O.lasso_path(x, y, [LAM], nthread=2)
    b = r["beta"][:, 0]
    assert r["niter"][0] == 339
    assert np.array_equal(b != 0, R.LASSO_PARADMM != 0)
    assert np.abs(b - R.LASSO_PARADMM).max() < 2e-6


def test_enet_readme_column(lasso_xy):
    x, y = lasso_xy
    r = O.lasso_path(x, y, [LAM], model="enet", alpha=0.5)
    b = r["beta"][:, 0]
    assert r["niter"][0] == 22
    assert np.array_equal(b != 0, R.ENET_ADMM != 0)
    assert np.abs(b - R.ENET_ADMM).max() < 5e-6


def test_lad_readme_column(lasso_xy):
    x, y = lasso_xy
    r = O.lad(x, y, intercept=False)
    assert r["niter"] == 443
    assert r["beta"][0] == 0.0
    assert np.abs(r["beta"][1:] - R.LAD_ADMM).max() < 1e-9
    assert np.abs(r["beta"][1:] - R.LAD_RQ).max() < 7e-3      # README.md:332 band vs quantreg


def test_bp_readme_range():
    d = np.load(os.path.join(G, "readme_bp_data.npz"))
    r = O.bp(d["x"], d["y"])
    diff = d["beta_true"] - r["beta"]
    assert r["niter"] == 72
    assert abs(diff.min() - R.BP_RANGE[0]) < 1e-10
    assert abs(diff.max() - R.BP_RANGE[1]) < 1e-10


def test_coarse_eigenvalue_is_coarse_and_deterministic(lasso_xy):
    x, y = lasso_xy
    xs = np.asfortranarray(x, dtype=np.float32).copy(order="F")
    ys = y.astype(np.float32)
    O.standardize_f32(xs, ys)
    Gm = O.gram_tn_f32(xs)
    ev, info = O.coarse_eig_f32(Gm)
    ev2, _ = O.coarse_eig_f32(Gm)
    exact = np.linalg.eigvalsh((np.tril(Gm) + np.tril(Gm, -1).T).astype(np.float64)).max()
    assert ev == ev2
    assert info["converged"] == 1 and info["nmatvec"] <= 13
    assert 0.85 * exact < ev <= exact * (1 + 1e-6)
