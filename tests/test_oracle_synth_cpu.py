"""The oracle's copy of the benchmark generator and its streamed full-size fit (CPU only).

The generator follows a fixed arithmetic recipe (oracle/synth_stream.hpp) so that the library's CUDA generator
produces the bit-identical design -- that equality is a GPU test (tests/test_gpu_kernels.py); here: the stream is
a sane N(mean, sd^2) sample, any row block can be produced independently, and the streamed fit (X never
resident, three generation passes) agrees with the in-memory oracle on the same matrix."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def O():
    from oracle import pyoracle
    pyoracle.use_openblas(4)
    return pyoracle


def test_generator_moments_and_blocks(O):
    X, y = O.synth_f32(40000, 64, row0=0, seed=123, mean_x=0.5, sd_x=2.0, nsig=10, noise=1.0)
    assert abs(float(X.mean()) - 0.5) < 0.01 and abs(float(X.std()) - 2.0) < 0.01
    z = (X[:, :8].astype(np.float64) - 0.5) / 2.0
    assert abs(float((z ** 3).mean())) < 0.03 and abs(float((z ** 4).mean()) - 3.0) < 0.08     # skewness, kurtosis
    c = np.corrcoef(X[:, :16].T)
    assert np.abs(c - np.eye(16)).max() < 0.03                                                 # independent columns
    # any row block, aligned to the 4-row counter groups or not, reproduces the same global rows
    for r0, nr in ((0, 7), (4, 64), (777, 1001), (39990, 10)):
        Xb, yb = O.synth_f32(nr, 64, row0=r0, seed=123, mean_x=0.5, sd_x=2.0, nsig=10, noise=1.0)
        assert np.array_equal(Xb, X[r0:r0 + nr]) and np.array_equal(yb, y[r0:r0 + nr])
    # y = X beta* + noise with beta* ~ U(0,1) on the first nsig columns
    b, *_ = np.linalg.lstsq(X[:, :10].astype(np.float64), (y - 0.0).astype(np.float64), rcond=None)
    assert (b > -0.05).all() and (b < 1.05).all()
    res = y - X[:, :10].astype(np.float64) @ b
    assert abs(res.std() - 1.0) < 0.02
    X2, _ = O.synth_f32(100, 64, seed=124)
    assert not np.array_equal(X2, X[:100])


def test_streamed_fit_matches_in_memory_oracle(O):
    n, p, nl = 30000, 96, 15
    X, y = O.synth_f32(n, p, nsig=12)
    r = O.tall_fit_synth(n, p, nsig=12, chunk_rows=4096, nlambda=nl, want_gram=True)
    o = O.lasso_path(X.astype(np.float64), y.astype(np.float64), nlambda=nl)
    assert np.allclose(r["lambda_"], o["lambda_"], rtol=1e-6)
    assert abs(r["rho"] - o["rho"]) < 1e-4 * o["rho"]
    # chunked column statistics / chunked SYRK differ from the whole-column sums in summation order only
    assert np.abs(r["beta"] - o["beta"]).max() < 2e-4
    assert abs(int(r["niter"].sum()) - int(o["niter"].sum())) <= max(3, 0.05 * int(o["niter"].sum()))
    Xs = np.asfortranarray(X.copy()); ys = y.copy()
    O.standardize_f32(Xs, ys)
    G = Xs.astype(np.float64).T @ Xs.astype(np.float64)
    # (float column norms: the whole-column and the chunked accumulation each carry ~1e-6 relative error)
    assert np.abs(np.tril(r["gram"]) - np.tril(G)).max() < 5e-6 * n
    assert np.allclose(r["xy"], Xs.astype(np.float64).T @ ys, rtol=0, atol=5e-6 * n)
    t = r["times"]
    assert t["total"] >= t["generate"] + t["gram"] > 0
