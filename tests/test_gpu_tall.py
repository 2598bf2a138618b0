"""GPU parity tests of the tall (n > p) lasso / elastic-net path: the CUDA library, called through
its C ABI (via the Python mirror of the R chain), against the CPU oracle on the same inputs and
against the reference's README vectors.

Tolerances (stated): float32 path, same (X, y, lambda, rho):
  * coefficients: max|dbeta| <= max(2e-4, 2 sqrt(p) eps_abs) * max(1, |beta|_inf) on the original scale.
    The stopping rule accepts any iterate with |r|_2 < sqrt(p) eps_abs + eps_rel |x|_2 (eps = 1e-5), so two
    runs that stop an iteration apart -- GPU and CPU differ in the summation order of the norms, of the
    K^-1 product and of the Gram matrix -- may differ by that much; measured over six problems
    (tests/tools/parity_modes.py, p = 200 .. 1024): up to 1.3e-4 with the TF32 Gram split, 2.2e-4 with the
    fp16 split (whose Gram matrix is the closer of the two to float64);
  * support identical except coordinates whose magnitude is below 1e-4 in either solution;
  * per-iteration scalars (eps, residuals) within 1e-3 relative over the first iterations;
  * iteration counts within +-2 per lambda or 3 % in total.
"""
import os

import numpy as np
import pytest

import readme_vectors as R

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAM = float(np.exp(-2))


@pytest.fixture(scope="module")
def O():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="module")
def A():
    import admm_b200
    admm_b200.device_info()
    return admm_b200


@pytest.fixture(scope="module")
def lasso_xy():
    d = np.load(os.path.join(G, "readme_lasso_data.npz"))
    return d["x"], d["y"]


def dense(beta):
    return np.asarray(beta.todense())


def assert_beta_close(b_gpu, b_cpu, tol=2e-4, band=2e-4):
    scale = max(1.0, float(np.abs(b_cpu).max()))
    assert np.abs(b_gpu - b_cpu).max() <= tol * scale, np.abs(b_gpu - b_cpu).max()
    mism = (b_gpu != 0) != (b_cpu != 0)
    if mism.any():
        assert max(np.abs(b_gpu[mism]).max(), np.abs(b_cpu[mism]).max()) < band


def test_readme_lasso(A, O, lasso_xy):
    x, y = lasso_xy
    f = A.admm_lasso(x, y).penalty(LAM).fit()
    o = O.lasso_path(x, y, [LAM])
    b = dense(f.beta)[:, 0]
    assert np.array_equal(b != 0, R.LASSO_ADMM != 0)                  # support bit-exact vs the reference's README
    assert np.abs(b - R.LASSO_ADMM).max() < 2e-5
    assert np.abs(b - o["beta"][:, 0]).max() < 5e-6
    assert abs(int(f.niter[0]) - int(o["niter"][0])) <= 1
    assert abs(f.info["rho"] - o["rho"]) < 1e-5 * o["rho"]


def test_readme_tall_columns_at_print_precision_with_the_rho_of_the_build_that_knitted_them(A, O, lasso_xy):
    """The README was knitted by a build whose Spectra call kept two Lanczos vectors (tests/test_oracle_golden.py); the library
    follows today's source (ncv = 3), so the rho of that build is handed over through opts(rho = ...): both tall columns
    of the README then come out of the CUDA library within 3e-6 / 4e-6 (the intercept, 5.36, carries 6 float32 ulps of
    the recover step's summation order; coefficients within 1.3e-6) against 1.2e-5 for the lasso column with the default rho."""
    x, y = lasso_xy
    with O.lanczos_ncv(2):
        rho_l = O.lasso_path(x, y, [LAM])["rho"]
        rho_e = O.lasso_path(x, y, [LAM], model="enet", alpha=0.5)["rho"]
    f = A.admm_lasso(x, y).penalty(LAM).opts(rho=rho_l).fit()
    bl = dense(f.beta)[:, 0]
    g = A.admm_enet(x, y).penalty(LAM, alpha=0.5).opts(rho=rho_e).fit()
    be = dense(g.beta)[:, 0]
    print("\n[readme] GPU with the README build's rho: lasso column %.2e (niter %d), enet column %.2e (niter %d)"
          % (np.abs(bl - R.LASSO_ADMM).max(), int(f.niter[0]), np.abs(be - R.ENET_ADMM).max(), int(g.niter[0])))
    assert np.abs(bl - R.LASSO_ADMM).max() < 5e-6 and np.array_equal(bl != 0, R.LASSO_ADMM != 0) and int(f.niter[0]) == 31
    assert np.abs(be - R.ENET_ADMM).max() < 5e-6 and np.array_equal(be != 0, R.ENET_ADMM != 0) and int(g.niter[0]) == 22
    f0 = A.admm_lasso(x, y).penalty(LAM).fit()
    assert np.abs(dense(f0.beta)[:, 0] - R.LASSO_ADMM).max() > 8e-6                 # today's source (ncv = 3)


def test_readme_enet(A, O, lasso_xy):
    x, y = lasso_xy
    f = A.admm_enet(x, y).penalty(LAM, alpha=0.5).fit()
    o = O.lasso_path(x, y, [LAM], model="enet", alpha=0.5)
    b = dense(f.beta)[:, 0]
    assert np.array_equal(b != 0, R.ENET_ADMM != 0)
    assert np.abs(b - R.ENET_ADMM).max() < 5e-6
    assert np.abs(b - o["beta"][:, 0]).max() < 5e-6
    assert abs(int(f.niter[0]) - int(o["niter"][0])) <= 1


def make_problem(n, p, seed, nsig=10, mean=0.0):
    rng = np.random.default_rng(seed)
    x = rng.normal(mean, 2.0, size=(n, p))
    b = np.zeros(p)
    b[:nsig] = rng.uniform(size=nsig)
    y = x @ b + rng.normal(size=n)
    return np.asfortranarray(x), y


@pytest.mark.parametrize("n,p,model,alpha", [(2000, 200, "lasso", 1.0), (1501, 137, "lasso", 1.0),
                                             (3000, 256, "enet", 0.5), (400, 50, "enet", 0.3)])
def test_path_matches_oracle(A, O, n, p, model, alpha):
    x, y = make_problem(n, p, seed=n + p)
    nl = 25
    if model == "lasso":
        f = A.admm_lasso(x, y).penalty(nlambda=nl).fit()
    else:
        f = A.admm_enet(x, y).penalty(nlambda=nl, alpha=alpha).fit()
    o = O.lasso_path(x, y, nlambda=nl, model=model, alpha=alpha)
    assert np.allclose(f.lambda_, o["lambda_"], rtol=1e-5)
    assert abs(f.info["rho"] - o["rho"]) < 1e-4 * o["rho"]
    bg, bc = dense(f.beta), o["beta"]
    tol = max(2e-4, 2.0 * np.sqrt(p) * 1e-5)
    for k in range(nl):
        assert_beta_close(bg[:, k], bc[:, k], tol=tol, band=tol)
    ng, nc = f.niter.astype(int), o["niter"].astype(int)
    # (5 % on the total: at the small-lambda end the oracle's own counts oscillate -- 37, 4, 49, 4, 42, 4, 41 measured for
    # enet n = 3000, p = 256 -- and shift with the summation order of the host BLAS, i.e. with the box's core count)
    assert abs(ng.sum() - nc.sum()) <= max(3, 0.05 * nc.sum()), (ng, nc)
    # Early in the path the counts are identical.  At the small-lambda end the warm-started
    # iterate sits at the float32 noise floor of the stopping rule (eps 1e-5 relative on float
    # vectors), so a last-ulp difference in a norm moves single lambdas by a few iterations in
    # either direction while the total stays put; report, bound loosely.
    head = max(5, nl // 2)
    assert np.abs(ng[:head] - nc[:head]).max() <= 2, (ng, nc)
    assert abs(ng[head:].sum() - nc[head:].sum()) <= max(4, 0.1 * nc[head:].sum()), (ng, nc)


@pytest.mark.parametrize("standardize,intercept", [(True, True), (True, False), (False, True), (False, False)])
def test_standardize_flags(A, O, standardize, intercept):
    x, y = make_problem(500, 40, seed=11, mean=1.2)
    y = y + 5.0
    lam = [0.5, 0.1, 0.02]
    f = A.admm_lasso(x, y, intercept=intercept, standardize=standardize).penalty(lam).fit()
    o = O.lasso_path(x, y, lam, standardize=standardize, intercept=intercept)
    bg, bc = dense(f.beta), o["beta"]
    for k in range(len(lam)):
        assert_beta_close(bg[:, k], bc[:, k], tol=2e-4)
    if not intercept:
        assert np.all(bg[0] == 0)


def test_iteration_trace_matches_oracle(A, O):
    from admm_b200 import _capi as K
    x, y = make_problem(1200, 96, seed=5)
    lam = [0.05]
    with K.trace(which=0, cap=4000) as tr:
        f = A.admm_lasso(x, y).penalty(lam).fit()
    o = O.lasso_path(x, y, lam, trace_lambda=0, trace_cap=4000)
    tg, tc = tr.rows, o["trace"][: int(o["niter"][0])]
    m = min(len(tg), len(tc), 15)
    assert m >= 5
    # eps_primal, resid_primal, eps_dual, resid_dual, rho
    assert np.allclose(tg[:m], tc[:m], rtol=1e-3, atol=1e-7), np.abs(tg[:m] / tc[:m] - 1).max()
    assert len(tg) == int(f.niter[0])


def test_input_dtypes_agree(A):
    import torch
    x, y = make_problem(800, 64, seed=3)
    f64 = A.admm_lasso(x, y).penalty(nlambda=5).fit()
    x32 = np.asfortranarray(x.astype(np.float32))
    y32 = y.astype(np.float32)
    f32 = A.admm_lasso(x32, y32).penalty(nlambda=5).fit()
    xd = torch.from_numpy(np.ascontiguousarray(x32.T)).cuda().t()      # column-major on the device
    yd = torch.from_numpy(y32).cuda()
    fd = A.admm_lasso(xd, yd).penalty(nlambda=5).fit()
    assert np.array_equal(f64.beta.toarray(), f32.beta.toarray())      # same float32 data after the narrowing copy
    assert np.array_equal(f32.beta.toarray(), fd.beta.toarray())
    assert np.array_equal(xd.t().cpu().numpy(), np.ascontiguousarray(x32.T))  # caller's device copy untouched


def test_pipelined_host_ingest_is_bit_identical(A, monkeypatch):
    """Host inputs with p >= 1024 are copied, standardised and folded into the Gram matrix panel by
    panel on two streams; the result must be bit-identical to the unpipelined device-input path."""
    import torch
    x, y = make_problem(2500, 1100, seed=8, nsig=15)
    x32 = np.asfortranarray(x.astype(np.float32))
    y32 = y.astype(np.float32)
    xd = torch.from_numpy(np.ascontiguousarray(x32.T)).cuda().t()
    yd = torch.from_numpy(y32).cuda()
    ref = A.admm_lasso(xd, yd).penalty(nlambda=6).fit()                 # device input: single pass
    monkeypatch.setenv("B200ADMM_PANEL_COLS", "256")                    # 5 panels
    f32 = A.admm_lasso(x32, y32).penalty(nlambda=6).fit()
    f64 = A.admm_lasso(x, y).penalty(nlambda=6).fit()
    monkeypatch.setenv("B200ADMM_PIPELINE", "0")
    f32_plain = A.admm_lasso(x32, y32).penalty(nlambda=6).fit()
    for f in (f32, f64, f32_plain):
        assert np.array_equal(f.beta.toarray(), ref.beta.toarray())
        assert np.array_equal(f.niter, ref.niter)
    # unaligned n (padded leading dimension, 2-D copies)
    x2, y2 = make_problem(2501, 1100, seed=9, nsig=15)
    monkeypatch.delenv("B200ADMM_PIPELINE")
    a = A.admm_lasso(x2, y2).penalty(nlambda=4).fit()
    monkeypatch.setenv("B200ADMM_PIPELINE", "0")
    b = A.admm_lasso(x2, y2).penalty(nlambda=4).fit()
    assert np.array_equal(a.beta.toarray(), b.beta.toarray())


def test_fp16_and_tf32_gram_paths_agree(A, O, monkeypatch):
    """DataStd flag 3 with p >= 256 builds the Gram matrix from pre-split fp16 operands; B200ADMM_GRAM=tf32
    forces the TF32 split.  Both are float32-accurate, so the fits agree within the stopping tolerance and
    both match the oracle; without scaling (standardize = FALSE) the TF32 kernel is used whatever the data."""
    x, y = make_problem(5000, 384, seed=17, nsig=12, mean=0.7)
    lam = [0.3, 0.05, 0.01]
    f16 = A.admm_lasso(x, y).penalty(lam).fit()
    monkeypatch.setenv("B200ADMM_GRAM", "tf32")
    f32 = A.admm_lasso(x, y).penalty(lam).fit()
    monkeypatch.delenv("B200ADMM_GRAM")
    o = O.lasso_path(x, y, lam)
    tol = max(2e-4, 2.0 * np.sqrt(384) * 1e-5)
    for k in range(len(lam)):
        assert_beta_close(dense(f16.beta)[:, k], o["beta"][:, k], tol=tol, band=tol)
        assert_beta_close(dense(f32.beta)[:, k], o["beta"][:, k], tol=tol, band=tol)
        assert_beta_close(dense(f16.beta)[:, k], dense(f32.beta)[:, k], tol=tol, band=tol)
    assert abs(f16.info["rho"] - f32.info["rho"]) < 1e-5 * f32.info["rho"]
    # any other DataStd flag keeps the TF32 split (here flag 1: scaled, not centred; CTA-pair TF32 kernel at p >= 256)
    g = A.admm_lasso(x, y, intercept=False).penalty(lam).fit()
    og = O.lasso_path(x, y, lam, intercept=False)
    for k in range(len(lam)):
        assert_beta_close(dense(g.beta)[:, k], og["beta"][:, k], tol=tol, band=tol)


def test_user_rho_and_nonconvergence(A, O):
    x, y = make_problem(600, 30, seed=9)
    f = A.admm_lasso(x, y).penalty([0.1]).opts(maxit=3, rho=50.0).fit()
    o = O.lasso_path(x, y, [0.1], maxit=3, rho=50.0)
    assert int(f.niter[0]) == 4 == int(o["niter"][0])                   # maxit + 1 (FADMMBase.h:264)
    assert np.abs(dense(f.beta)[:, 0] - o["beta"][:, 0]).max() < 1e-5


def test_default_lambda_min_ratio_is_resolved_by_the_library(A):
    """A default lambda_min_ratio goes to the C ABI as 0 and is resolved there from the GLOBAL shape (a rank of a
    row-sharded run can hold fewer rows than columns; tools/mgpu_bisect.py): same grid as the explicit R default."""
    rng = np.random.default_rng(5)
    for (n, p, ratio) in ((400, 30, 1e-4), (30, 90, 0.01)):
        x = np.asfortranarray(rng.normal(size=(n, p)))
        y = x[:, :3] @ np.array([1.0, -0.5, 0.25]) + 0.1 * rng.normal(size=n)
        fd = A.admm_lasso(x, y).penalty(nlambda=7).fit()
        fe = A.admm_lasso(x, y).penalty(nlambda=7, lambda_min_ratio=ratio).fit()
        assert np.array_equal(fd.lambda_, fe.lambda_)
        assert abs(fd.lambda_[-1] / fd.lambda_[0] - ratio) < 1e-9 * ratio + 1e-12


def test_too_few_variables_is_an_error(A):
    from admm_b200 import B200AdmmError
    x, y = make_problem(50, 2, seed=1, nsig=1)
    with pytest.raises(B200AdmmError) as e:
        A.admm_lasso(x, y).penalty([0.1]).fit()
    assert e.value.code == -5


def test_kkt_standardised_path_at_scale(A):
    """Size-independent property at a size that exercises the whole production path of the headline config
    (DataStd flag 3 -> fused standardise / X'y / fp16 operand split -> CTA-pair Gram with full rounds and a
    K-split tail (153 tiles over 74 pairs) -> graph-replayed factorisation -> persistent path kernel): the
    returned coefficients satisfy the lasso optimality conditions of the standardised problem,
    |x_j'(y - b0 - X b)| / (n sd_j) <= lambda with equality and matching sign on the support."""
    import torch
    n, p = 200_000, 4352
    g = torch.Generator(device="cuda").manual_seed(77)
    X = torch.randn((p, n), device="cuda", dtype=torch.float32, generator=g) * 2.0 + 0.5       # column-major n x p
    bt = torch.zeros(p, device="cuda", dtype=torch.float32)
    bt[:40] = torch.rand(40, device="cuda", generator=g) + 0.2
    y = X.t() @ bt + torch.randn(n, device="cuda", generator=g) + 3.0
    f0 = A.admm_lasso(X.t(), y).penalty(nlambda=2).opts(maxit=1).fit()
    lam = [0.2 * float(f0.lambda_[0]), 0.02 * float(f0.lambda_[0])]
    f = A.admm_lasso(X.t(), y).penalty(lam).opts(eps_abs=1e-7, eps_rel=1e-7).fit()
    B = torch.from_numpy(dense(f.beta)).cuda()                           # (p + 1) x 2, float64
    Xd = X.double()
    sd = Xd.std(dim=1, unbiased=False)
    for k, lk in enumerate(lam):
        r = y.double() - B[0, k] - Xd.t() @ B[1:, k]
        assert abs(float(r.mean())) < 1e-5
        grad = (Xd @ r) / n / sd
        s = B[1:, k] != 0
        assert int(s.sum()) >= 30                      # the weakest of the 40 signals may stay below a large lambda
        assert float(grad.abs().max()) <= lk * (1 + 2e-3)
        assert torch.allclose(grad[s], lk * torch.sign(B[1:, k][s]), rtol=5e-3, atol=0)


def test_kkt_at_moderate_size(A):
    """Size-independent property: the solution satisfies the lasso optimality conditions
    |X_j'(y - X b)| <= n*lambda (+tol), with equality and matching sign on the support."""
    n, p = 20000, 500
    x, y = make_problem(n, p, seed=21, nsig=20)
    lam = 0.05
    f = A.admm_lasso(x, y, standardize=False, intercept=False).penalty([lam]).opts(eps_abs=1e-7, eps_rel=1e-7).fit()
    b = dense(f.beta)[1:, 0]
    g = x.T @ (y - x @ b) / n
    assert np.abs(g).max() <= lam * (1 + 2e-3)
    s = b != 0
    assert s.sum() >= 15
    assert np.allclose(g[s], lam * np.sign(b[s]), rtol=5e-3)
