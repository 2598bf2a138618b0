"""The reference-side Rcpp glue (r-pkg/src/b200_glue.cpp) EXECUTED against the CUDA library on the GPU: compiled with the
functional <Rcpp.h> stand-in of tests/stubs/ (R is not in the image), linked to admm_b200/libb200admm.so the way
r-pkg/src/Makevars does, and called with the arguments the reference's R front end passes to .Call (tests/rstub.py).
The returned R objects -- List(lambda, beta = dgCMatrix, niter) etc. -- must equal what the Python mirror of the R chain
gets from the same library, and the README's worked examples must come out of the R entry points."""
import os

import numpy as np
import pytest

import readme_vectors as R
import rstub

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAM = float(np.exp(-2))


@pytest.fixture(scope="module")
def A():
    import admm_b200
    admm_b200.device_info()
    return admm_b200


@pytest.fixture(scope="module")
def glue(tmp_path_factory):
    return rstub.build(tmp_path_factory.mktemp("rglue_cuda"), fake=False)


@pytest.fixture(scope="module")
def lasso_xy():
    d = np.load(os.path.join(G, "readme_lasso_data.npz"))
    return d["x"], d["y"]


def dense(beta):
    return np.asarray(beta.todense())


def test_readme_examples_through_the_r_entry_points(glue, lasso_xy):
    """README.md:52-88, :95-122, :130-160: the printed columns from .Call("admm_lasso" / "admm_parlasso" / "admm_enet" / "admm_lad")."""
    x, y = lasso_xy
    r = rstub.r_admm_lasso_fit(glue, x, y, lam=[LAM])
    b = rstub.dgc_to_dense(r["beta"])[:, 0]
    assert list(r) == ["lambda", "beta", "niter"] and r["lambda"][0] == LAM
    assert np.abs(b - R.LASSO_ADMM).max() < 2e-5 and np.array_equal(b != 0, R.LASSO_ADMM != 0)
    r = rstub.r_admm_lasso_fit(glue, x, y, lam=[LAM], nthread=2)
    b = rstub.dgc_to_dense(r["beta"])[:, 0]
    assert np.abs(b - R.LASSO_PARADMM).max() < 5e-6 and np.array_equal(b != 0, R.LASSO_PARADMM != 0)
    r = rstub.r_admm_enet_fit(glue, x, y, lam=[LAM], alpha=0.5)
    b = rstub.dgc_to_dense(r["beta"])[:, 0]
    assert np.abs(b - R.ENET_ADMM).max() < 5e-6 and np.array_equal(b != 0, R.ENET_ADMM != 0)
    r = rstub.r_admm_lad_fit(glue, x, y, intercept=False)
    assert r["beta"].shape == (21,) and r["beta"][0] == 0.0 and np.abs(r["beta"][1:] - R.LAD_ADMM).max() < 1e-7


def test_r_objects_equal_the_python_mirror(glue, A, lasso_xy):
    """Same library, two front ends: the dgCMatrix / lambda / niter the glue builds are the Python mirror's, bit for bit."""
    x, y = lasso_xy
    f = A.admm_lasso(x, y).penalty(nlambda=25).fit()
    r = rstub.r_admm_lasso_fit(glue, x, y, nlambda=25)
    assert np.array_equal(r["lambda"], np.asarray(f.lambda_)) and np.array_equal(r["niter"], np.asarray(f.niter))
    assert np.array_equal(rstub.dgc_to_dense(r["beta"]), dense(f.beta))
    assert np.array_equal(r["beta"]["Dim"], [21, 25]) and r["beta"]["p"][-1] == len(r["beta"]["x"]) == len(r["beta"]["i"])
    assert np.all(r["beta"]["i"][r["beta"]["p"][:-1]] == 0)                      # the intercept row is always stored
    # wide branch, explicit lambdas given in increasing order (R sorts them decreasingly), options forwarded
    rng = np.random.default_rng(3)
    xw = rng.normal(size=(40, 120))
    yw = xw[:, :5] @ np.ones(5) + 0.1 * rng.normal(size=40)
    lam = [0.05, 0.2, 0.1]
    f = A.admm_lasso(xw, yw).penalty(lam).opts(maxit=500, eps_abs=1e-6, eps_rel=1e-6).fit()
    r = rstub.r_admm_lasso_fit(glue, xw, yw, lam=lam, maxit=500, eps_abs=1e-6, eps_rel=1e-6)
    assert np.array_equal(r["lambda"], [0.2, 0.1, 0.05]) and np.array_equal(r["niter"], np.asarray(f.niter))
    assert np.array_equal(rstub.dgc_to_dense(r["beta"]), dense(f.beta))
    # basis pursuit: List(beta = dgCMatrix p x 1, niter) without an intercept row
    d = np.load(os.path.join(G, "readme_bp_data.npz"))
    f = A.admm_bp(d["x"], d["y"]).fit()
    r = rstub.r_admm_bp_fit(glue, d["x"], d["y"])
    assert list(r) == ["beta", "niter"] and int(r["niter"][0]) == int(f.niter)
    b = rstub.dgc_to_dense(r["beta"])
    assert b.shape == (100, 1) and np.array_equal(b, dense(f.beta))
    diff = d["beta_true"] - b[:, 0]
    assert abs(diff.min() - R.BP_RANGE[0]) < 1e-7 and abs(diff.max() - R.BP_RANGE[1]) < 1e-7    # README.md:180-182


def test_library_errors_surface_as_r_errors(glue):
    x = np.random.default_rng(0).normal(size=(30, 40))
    with pytest.raises(rstub.RError):
        rstub.r_admm_lasso_fit(glue, x, np.zeros(30), lam=[0.1], nthread=3, maxit=-5)   # rejected by the library, not by a crash
