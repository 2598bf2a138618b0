"""Host-side mirror of the reference's R front-end over the C ABI.

The reference exposes R5 RefClasses built by admm_lasso() / admm_enet() / admm_lad() / admm_bp()
with chainable $penalty() / $parallel() / $opts() and a final $fit() that does the .Call
(/root/reference/R/30_admm_lasso.R:26-160, R/40_admm_enet.R:16-63, R/20_admm_lad.R:16-68,
R/10_admm_bp.R:24-119).  R is not available in this image, so the same chain -- same names,
defaults, argument meaning and stop() conditions -- is restated here in Python; each fit() makes
exactly the call the R method makes, against libb200admm.so instead of ADMM.so.

    fit = admm_lasso(x, y).penalty(nlambda=20).opts(eps_rel=1e-6).fit()
    fit.lambda_, fit.beta ((p+1) x nlambda scipy CSC, intercept in row 0), fit.niter
"""
import ctypes as C

import numpy as np

from . import _capi as K


def _shape(x):
    return tuple(x.shape)


def _len(y):
    return int(y.shape[0]) if hasattr(y, "shape") else len(y)


class ADMM_Lasso_fit:
    """ADMM_Lasso_fit RefClass (R/30_admm_lasso.R:18-22): lambda, beta (dgCMatrix), niter."""
    title = "ADMM Lasso fitting result"

    def __init__(self, lambda_, beta, niter, info=None):
        self.lambda_ = lambda_
        self.beta = beta
        self.niter = niter
        self.info = info or {}

    def __repr__(self):
        return "%s\n\n$lambda: <%d> vector\n$beta: <%d x %d> sparse matrix\n$niter: <%d> vector\n" % (
            self.title, len(self.lambda_), self.beta.shape[0], self.beta.shape[1], len(self.niter))


class ADMM_Enet_fit(ADMM_Lasso_fit):
    title = "ADMM Elastic Net fitting result"


class ADMM_Lasso:
    """ADMM_Lasso RefClass (R/30_admm_lasso.R:26-160)."""

    def __init__(self, x, y, intercept=True, standardize=True):
        if _shape(x)[0] != _len(y):
            raise ValueError("nrow(x) should be equal to length(y)")
        self.x = x
        self.y = y
        self.intercept = bool(intercept)
        self.standardize = bool(standardize)
        self.lambda_ = np.zeros(0)
        self.nlambda = 100
        n, p = _shape(x)
        self.lambda_min_ratio = 0.01 if n < p else 0.0001
        self._lmr_is_default = True
        self.nthread = 1
        self.maxit = 10000
        self.eps_abs = 1e-5
        self.eps_rel = 1e-5
        self.rho = -1.0

    # $penalty(lambda = NULL, nlambda = 100, lambda_min_ratio)
    def penalty(self, lambda_=None, nlambda=100, lambda_min_ratio=None):
        lam = np.sort(np.atleast_1d(np.asarray([] if lambda_ is None else lambda_, dtype=np.float64)))[::-1].copy()
        if np.any(lam <= 0):
            raise ValueError("lambda must be positive")
        if nlambda <= 0:
            raise ValueError("nlambda must be a positive integer")
        n, p = _shape(self.x)
        lmr = (0.01 if n < p else 0.0001) if lambda_min_ratio is None else float(lambda_min_ratio)
        if lmr >= 1 or lmr <= 0:
            raise ValueError("lambda_min_ratio must be within (0, 1)")
        # a default is resolved by the library from the GLOBAL row count (a rank of a row-sharded run may hold fewer
        # rows than columns: measured on 8 ranks, 500 rows x 1100 columns each, the local default was 0.01 against 1e-4)
        self._lmr_is_default = lambda_min_ratio is None
        self.lambda_ = lam
        self.nlambda = int(nlambda)
        self.lambda_min_ratio = lmr
        return self

    # $parallel(nthread = 2)
    def parallel(self, nthread=2):
        nt = int(nthread)
        if nt < 1:
            nt = 1
        if nt >= _shape(self.x)[1] / 5:
            raise ValueError("nthread cannot exceed ncol(x)/5")
        self.nthread = nt
        return self

    # $opts(maxit = 10000, eps_abs = 1e-5, eps_rel = 1e-5, rho = NULL)
    def opts(self, maxit=10000, eps_abs=1e-5, eps_rel=1e-5, rho=None):
        if maxit <= 0:
            raise ValueError("maxit should be positive")
        if eps_abs < 0 or eps_rel < 0:
            raise ValueError("eps_abs and eps_rel should be nonnegative")
        if rho is not None and rho <= 0:
            raise ValueError("rho should be positive")
        self.maxit = int(maxit)
        self.eps_abs = float(eps_abs)
        self.eps_rel = float(eps_rel)
        self.rho = -1.0 if rho is None else float(rho)
        return self

    def _opts(self):
        return K.Opts(self.maxit, self.eps_abs, self.eps_rel, self.rho)

    def _lmr_arg(self):
        # 0 -> the library takes the reference's default from the global shape
        return 0.0 if self._lmr_is_default else self.lambda_min_ratio

    def _lambda_args(self):
        lam = np.ascontiguousarray(self.lambda_, dtype=np.float64)
        return lam, (lam.ctypes.data if lam.size else None), int(lam.size)

    fit_class = ADMM_Lasso_fit

    # $fit(): .Call("admm_lasso", ...) or .Call("admm_parlasso", ..., nthread, ...)
    def fit(self):
        d, keep = K.make_data(self.x, self.y)
        lam, lam_ptr, nlam = self._lambda_args()
        o = self._opts()
        P = K.Path()
        L = K.lib()
        if self.nthread <= 1:
            rc = L.b200admm_lasso(C.byref(d), lam_ptr, nlam, self.nlambda, self._lmr_arg(),
                                  int(self.standardize), int(self.intercept), C.byref(o), C.byref(P))
        else:
            rc = L.b200admm_parlasso(C.byref(d), lam_ptr, nlam, self.nlambda, self._lmr_arg(),
                                     int(self.standardize), int(self.intercept), self.nthread, C.byref(o), C.byref(P))
        K.check(rc)
        del keep
        return self.fit_class(*K.path_to_python(P))


class ADMM_Enet(ADMM_Lasso):
    """ADMM_Enet RefClass (R/40_admm_enet.R:16-63); $parallel() is inherited but fit() ignores it."""
    fit_class = ADMM_Enet_fit

    def __init__(self, x, y, intercept=True, standardize=True):
        super().__init__(x, y, intercept, standardize)
        self.alpha = 1.0

    def penalty(self, lambda_=None, nlambda=100, lambda_min_ratio=None, alpha=1):
        if alpha < 0 or alpha > 1:
            raise ValueError("alpha must be within [0,1]")
        self.alpha = float(alpha)
        return super().penalty(lambda_, nlambda, lambda_min_ratio)

    def fit(self):
        d, keep = K.make_data(self.x, self.y)
        lam, lam_ptr, nlam = self._lambda_args()
        o = self._opts()
        P = K.Path()
        rc = K.lib().b200admm_enet(C.byref(d), lam_ptr, nlam, self.nlambda, self._lmr_arg(),
                                   int(self.standardize), int(self.intercept), self.alpha, C.byref(o), C.byref(P))
        K.check(rc)
        del keep
        return self.fit_class(*K.path_to_python(P))


class ADMM_LAD_fit:
    """ADMM_LAD_fit (R/20_admm_lad.R:9-12): beta = c(intercept, coefficients), niter."""

    def __init__(self, beta, niter, info=None):
        self.beta = beta
        self.niter = niter
        self.info = info or {}


class ADMM_LAD:
    """ADMM_LAD RefClass (R/20_admm_lad.R:16-68)."""

    def __init__(self, x, y, intercept=True):
        n, p = _shape(x)
        if n <= p:
            raise ValueError("nrow(x) must be greater than ncol(x)")
        if n != _len(y):
            raise ValueError("nrow(x) should be equal to length(y)")
        self.x = x
        self.y = y
        self.maxit = 10000
        self.eps_abs = 1e-4
        self.eps_rel = 1e-4
        self.rho = 1.0
        self.intercept = bool(intercept)

    def opts(self, maxit=10000, eps_abs=1e-4, eps_rel=1e-4, rho=1.0):
        if maxit <= 0:
            raise ValueError("maxit should be positive")
        if eps_abs < 0 or eps_rel < 0:
            raise ValueError("eps_abs and eps_rel should be nonnegative")
        if rho <= 0:
            raise ValueError("rho should be positive")
        self.maxit = int(maxit)
        self.eps_abs = float(eps_abs)
        self.eps_rel = float(eps_rel)
        self.rho = float(rho)
        return self

    def fit(self):
        d, keep = K.make_data(self.x, self.y, want64=True)
        o = K.Opts(self.maxit, self.eps_abs, self.eps_rel, self.rho)
        D = K.Dense()
        K.check(K.lib().b200admm_lad(C.byref(d), int(self.intercept), C.byref(o), C.byref(D)))
        del keep
        beta = np.ctypeslib.as_array(D.beta, shape=(int(D.len),)).copy()
        info = dict(rho=D.rho, timing=D.t.as_dict())
        niter = int(D.niter)
        K.lib().b200admm_free_dense(C.byref(D))
        return ADMM_LAD_fit(beta, niter, info)


class ADMM_BP_fit:
    """ADMM_BP_fit (R/10_admm_bp.R:13-16): beta = p x 1 dgCMatrix, niter."""

    def __init__(self, beta, niter, info=None):
        self.beta = beta
        self.niter = niter
        self.info = info or {}


class ADMM_BP:
    """ADMM_BP RefClass (R/10_admm_bp.R:24-119)."""

    def __init__(self, x, y):
        n, p = _shape(x)
        if n >= p:
            raise ValueError("ncol(x) must be greater than nrow(x)")
        if n != _len(y):
            raise ValueError("nrow(x) should be equal to length(y)")
        self.x = x
        self.y = y
        self.nthread = 1
        self.maxit = 10000
        self.eps_abs = 1e-4
        self.eps_rel = 1e-4
        self.rho = 1.0

    def parallel(self, nthread=2):
        self.nthread = max(1, int(nthread))
        return self

    def opts(self, maxit=10000, eps_abs=1e-4, eps_rel=1e-4, rho=1):
        if maxit <= 0:
            raise ValueError("maxit should be positive")
        if eps_abs < 0 or eps_rel < 0:
            raise ValueError("eps_abs and eps_rel should be nonnegative")
        if rho <= 0:
            raise ValueError("rho should be positive")
        self.maxit = int(maxit)
        self.eps_abs = float(eps_abs)
        self.eps_rel = float(eps_rel)
        self.rho = float(rho)
        return self

    def fit(self):
        if self.nthread > 1:
            # the reference calls admm_parbp, whose native code is not part of the package
            # (src/TODO/ParBP.cppp is not compiled) -> R raises an error here as well
            raise RuntimeError('C symbol name "admm_parbp" not in DLL for package "ADMM"')
        d, keep = K.make_data(self.x, self.y, want64=True)
        o = K.Opts(self.maxit, self.eps_abs, self.eps_rel, self.rho)
        P = K.Path()
        K.check(K.lib().b200admm_bp(C.byref(d), C.byref(o), C.byref(P)))
        del keep
        lam, beta, niter, info = K.path_to_python(P)
        return ADMM_BP_fit(beta, int(niter[0]) if len(niter) else 0, info)


# constructors with the reference's exported names (NAMESPACE:9-13)
def admm_lasso(x, y, intercept=True, standardize=True):
    return ADMM_Lasso(x, y, intercept, standardize)


def admm_enet(x, y, intercept=True, standardize=True):
    return ADMM_Enet(x, y, intercept, standardize)


def admm_lad(x, y, intercept=True):
    return ADMM_LAD(x, y, intercept)


def admm_bp(x, y):
    return ADMM_BP(x, y)
