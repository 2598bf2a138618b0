"""ctypes front-end of the CPU oracle (oracle/admm_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs -- never
from admm_b200/.  Builds oracle/_build/liboracle.so on first use if g++ is available.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
_lib = None

c_i64 = C.c_longlong
c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("admm_oracle.cpp", "linalg.hpp", "lanczos.hpp", "synth_stream.hpp")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.check_call(["make", "-C", HERE], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_coarse_eig_f32.restype = C.c_float
    return _lib


def find_openblas():
    try:
        import scipy
    except Exception:
        return None
    d = os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs")
    c = glob.glob(os.path.join(d, "libscipy_openblas*.so"))
    return os.path.abspath(c[0]) if c else None


def use_openblas(threads=0):
    """Route the oracle's dense kernels through OpenBLAS (the reference's recommended build).
    Returns the BLAS thread count in effect, or 0 if no BLAS could be loaded."""
    p = find_openblas()
    if not p:
        return 0
    if lib().oracle_load_blas(p.encode()) != 0:
        return 0
    return lib().oracle_blas_threads(int(threads))


def omp_threads(n=0):
    return lib().oracle_omp_threads(int(n))


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _fp(a):
    return a.ctypes.data_as(c_fp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


def lasso_path(x, y, lambdas=None, nlambda=100, lambda_min_ratio=None, standardize=True, intercept=True,
               model="lasso", alpha=1.0, nthread=1, maxit=10000, eps_abs=1e-5, eps_rel=1e-5, rho=-1.0,
               trace_lambda=None, trace_cap=0):
    """admm_lasso / admm_enet / admm_parlasso of the reference, restated on the CPU.
    Returns dict(lambda, beta[(p+1) x nl], niter, aux, trace)."""
    x = np.asfortranarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n, p = x.shape
    if lambda_min_ratio is None:
        lambda_min_ratio = 0.01 if n < p else 1e-4
    if lambdas is None or len(lambdas) == 0:
        lam_in, nin, nl = None, 0, int(nlambda)
    else:
        lam_in = np.sort(np.asarray(lambdas, dtype=np.float64))[::-1].copy()
        nin = nl = len(lam_in)
    lam_out = np.zeros(nl)
    beta = np.zeros((p + 1, nl), order="F")
    niter = np.zeros(nl, dtype=np.int32)
    aux = np.zeros(8)
    trace = np.zeros((trace_cap, 5)) if trace_cap > 0 else None
    rc = lib().oracle_lasso_path(
        _dp(x), _dp(y), c_i64(n), c_i64(p), C.c_int(1 if model == "enet" else 0), C.c_double(alpha),
        _dp(lam_in), C.c_int(nin), C.c_int(nl), C.c_double(lambda_min_ratio),
        C.c_int(int(standardize)), C.c_int(int(intercept)), C.c_int(int(nthread)),
        C.c_int(int(maxit)), C.c_double(eps_abs), C.c_double(eps_rel), C.c_double(rho),
        _dp(lam_out), _dp(beta), _ip(niter), _dp(aux),
        _dp(trace), C.c_int(trace_cap), C.c_int(-1 if trace_lambda is None else int(trace_lambda)))
    if rc != 0:
        raise RuntimeError(f"oracle_lasso_path failed rc={rc}")
    return dict(lambda_=lam_out, beta=beta, niter=niter, aux=aux, trace=trace,
                rho=aux[0], eig=aux[1], lambda0=aux[2], scaleY=aux[3], meanY=aux[4])


class lanczos_ncv:
    """with lanczos_ncv(2): ...  -- the Spectra call behind the default rho / gamma run with another ncv (README forensics only;
    3 is the reference's source as it stands)."""

    def __init__(self, ncv):
        self.ncv = int(ncv)

    def __enter__(self):
        lib().oracle_set_lanczos_ncv.restype = C.c_int
        self.old = lib().oracle_set_lanczos_ncv(C.c_int(self.ncv))
        return self

    def __exit__(self, *exc):
        lib().oracle_set_lanczos_ncv(C.c_int(self.old))
        return False


def lad(x, y, intercept=True, maxit=10000, eps_abs=1e-4, eps_rel=1e-4, rho=1.0, trace_cap=0):
    x = np.asfortranarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n, p = x.shape
    beta = np.zeros(p + 1)
    niter = np.zeros(1, dtype=np.int32)
    trace = np.zeros((trace_cap, 5)) if trace_cap > 0 else None
    rc = lib().oracle_lad(_dp(x), _dp(y), c_i64(n), c_i64(p), C.c_int(int(intercept)), C.c_int(int(maxit)),
                          C.c_double(eps_abs), C.c_double(eps_rel), C.c_double(rho),
                          _dp(beta), _ip(niter), _dp(trace), C.c_int(trace_cap))
    if rc != 0:
        raise RuntimeError(f"oracle_lad failed rc={rc}")
    return dict(beta=beta, niter=int(niter[0]), trace=trace)


def bp(x, y, maxit=10000, eps_abs=1e-4, eps_rel=1e-4, rho=1.0, trace_cap=0):
    x = np.asfortranarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n, p = x.shape
    beta = np.zeros(p)
    niter = np.zeros(1, dtype=np.int32)
    trace = np.zeros((trace_cap, 5)) if trace_cap > 0 else None
    rc = lib().oracle_bp(_dp(x), _dp(y), c_i64(n), c_i64(p), C.c_int(int(maxit)),
                         C.c_double(eps_abs), C.c_double(eps_rel), C.c_double(rho),
                         _dp(beta), _ip(niter), _dp(trace), C.c_int(trace_cap))
    if rc != 0:
        raise RuntimeError(f"oracle_bp failed rc={rc}")
    return dict(beta=beta, niter=int(niter[0]), trace=trace)


def standardize_f32(X, Y, standardize=True, intercept=True):
    """In-place DataStd<float> on Fortran-ordered float32 X (n x p) and Y.  Returns stats."""
    assert X.dtype == np.float32 and X.flags.f_contiguous and Y.dtype == np.float32
    n, p = X.shape
    meanX = np.zeros(p, dtype=np.float32)
    scaleX = np.ones(p, dtype=np.float32)
    ys = np.zeros(2, dtype=np.float32)
    lib().oracle_standardize_f32(_fp(X), _fp(Y), c_i64(n), c_i64(p), C.c_int(int(standardize)),
                                 C.c_int(int(intercept)), _fp(meanX), _fp(scaleX), _fp(ys))
    return dict(meanX=meanX, scaleX=scaleX, meanY=float(ys[0]), scaleY=float(ys[1]))


def coarse_eig_f32(S):
    """Spectra-emulating coarse lambda_max of a symmetric float32 matrix (lower triangle read)."""
    S = np.asfortranarray(S, dtype=np.float32)
    n = S.shape[0]
    info = np.zeros(3, dtype=np.int32)
    ev = lib().oracle_coarse_eig_f32(_fp(S), c_i64(n), _ip(info))
    return float(ev), dict(nmatvec=int(info[0]), nrestart=int(info[1]), converged=int(info[2]))


def gram_tn_f32(X):
    X = np.asfortranarray(X, dtype=np.float32)
    n, p = X.shape
    G = np.zeros((p, p), dtype=np.float32, order="F")
    lib().oracle_gram_tn_f32(_fp(X), c_i64(n), c_i64(p), _fp(G))
    return G


def tall_path_from_gram(G, XY, lambda_internal, enet=False, alpha=1.0, maxit=10000, eps_abs=1e-5,
                        eps_rel=1e-5, rho=-1.0, trace_lambda=None, trace_cap=0):
    """Tall lasso/enet lambda-path on the standardised scale, starting from lower(X'X) and X'y."""
    G = np.asfortranarray(G, dtype=np.float32)
    XY = np.ascontiguousarray(XY, dtype=np.float32)
    lam = np.ascontiguousarray(lambda_internal, dtype=np.float64)
    p = G.shape[0]
    nl = len(lam)
    z = np.zeros((nl, p), dtype=np.float32)
    niter = np.zeros(nl, dtype=np.int32)
    aux = np.zeros(4)
    trace = np.zeros((trace_cap, 5)) if trace_cap > 0 else None
    rc = lib().oracle_tall_path_from_gram(
        _fp(G), _fp(XY), c_i64(p), C.c_int(int(enet)), C.c_double(alpha), _dp(lam), C.c_int(nl),
        C.c_int(int(maxit)), C.c_double(eps_abs), C.c_double(eps_rel), C.c_double(rho),
        _fp(z), _ip(niter), _dp(aux), _dp(trace), C.c_int(trace_cap),
        C.c_int(-1 if trace_lambda is None else int(trace_lambda)))
    if rc != 0:
        raise RuntimeError(f"oracle_tall_path_from_gram failed rc={rc}")
    return dict(z=z, niter=niter, rho=aux[0], eig=aux[1], nmatvec=int(aux[2]), setup_s=aux[3], trace=trace)


def synth_f32(nrows, p, row0=0, seed=123, mean_x=0.0, sd_x=2.0, nsig=100, noise=1.0):
    """Rows [row0, row0 + nrows) of the synthetic benchmark design, bit-identical to the library's CUDA
    generator (b200admm_synth_f32).  Returns (X float32 Fortran-ordered nrows x p, y float32)."""
    X = np.empty((nrows, p), dtype=np.float32, order="F")
    y = np.empty(nrows, dtype=np.float32)
    lib().oracle_synth_f32(_fp(X), _fp(y), c_i64(nrows), c_i64(p), c_i64(row0), C.c_ulonglong(seed),
                           C.c_float(mean_x), C.c_float(sd_x), C.c_int(nsig), C.c_float(noise))
    return X, y


def tall_fit_synth(n, p, seed=123, mean_x=0.0, sd_x=2.0, nsig=100, noise=1.0, chunk_rows=32768, nlambda=100,
                   enet=False, alpha=1.0, lambda_frac=0.0, lambda_min_ratio=1e-4, maxit=10000, eps_abs=1e-5, eps_rel=1e-5, rho=-1.0, want_gram=False):
    """admm_lasso(x, y)$penalty(nlambda)$fit() of the reference on the full synthetic design, X streamed in row
    chunks (never resident).  Returns lambda, beta[(p+1) x nl], niter, times (s) and rho/eig/lambda0/scaleY/meanY."""
    if lambda_frac > 0:
        nlambda = 1
    lam = np.zeros(nlambda)
    beta = np.zeros((p + 1, nlambda), order="F")
    niter = np.zeros(nlambda, dtype=np.int32)
    times = np.zeros(8)
    aux = np.zeros(6)
    G = np.zeros((p, p), dtype=np.float32, order="F") if want_gram else None
    xy = np.zeros(p, dtype=np.float32) if want_gram else None
    rc = lib().oracle_tall_fit_synth(
        c_i64(n), c_i64(p), C.c_ulonglong(seed), C.c_float(mean_x), C.c_float(sd_x), C.c_int(nsig), C.c_float(noise),
        c_i64(chunk_rows), C.c_int(int(enet)), C.c_double(alpha), C.c_double(lambda_frac),
        C.c_int(nlambda), C.c_double(lambda_min_ratio), C.c_int(maxit), C.c_double(eps_abs),
        C.c_double(eps_rel), C.c_double(rho), _dp(lam), _dp(beta), _ip(niter), _dp(times), _dp(aux), _fp(G), _fp(xy))
    if rc != 0:
        raise RuntimeError(f"oracle_tall_fit_synth failed rc={rc}")
    t = dict(zip(("generate", "standardize", "gram", "lanczos_cholesky", "iterations", "total"), times[:6].tolist()))
    return dict(lambda_=lam, beta=beta, niter=niter, times=t, rho=aux[0], eig=aux[1], lambda0=aux[2], scaleY=aux[3],
                meanY=aux[4], gram=G, xy=xy)
