"""Execute the reference-side Rcpp glue (r-pkg/src/b200_glue.cpp) without R: the glue is compiled against the functional
<Rcpp.h> stand-in in tests/stubs/ and linked either to a recording fake of the library (CPU tests) or to the CUDA library
(GPU tests); `dot_call` plays R's .Call() with R's argument types (numeric matrix / vector, integer, logical, named list)
and converts the returned object (list, dgCMatrix, integer / numeric vectors, or an R error) to Python values."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "tests", "stubs")
NIL, LGL, INT, REAL, LIST, S4, ERROR = 0, 10, 13, 14, 19, 25, 99


class RError(RuntimeError):
    """What END_RCPP hands to R when the glue calls stop()."""


def build(outdir, fake):
    """Compile glue + host helpers into one shared object; `fake` links the recording fake instead of libb200admm.so."""
    so = os.path.join(str(outdir), "librglue_%s.so" % ("fake" if fake else "cuda"))
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-shared", "-fPIC", "-I", STUBS, "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "r-pkg", "src", "b200_glue.cpp"), os.path.join(STUBS, "rstub_host.cpp"), "-o", so]
    if fake:
        obj = os.path.join(str(outdir), "fake_b200admm.o")
        subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-fPIC", "-c", "-I", os.path.join(ROOT, "include"),
                        os.path.join(STUBS, "fake_b200admm.c"), "-o", obj], check=True, capture_output=True, text=True)
        cmd.insert(-2, obj)
    else:
        libdir = os.path.join(ROOT, "admm_b200")
        cmd += ["-L", libdir, "-lb200admm", "-Wl,-rpath," + libdir]      # what r-pkg/src/Makevars does with PKG_LIBS
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    return Glue(so)


class Glue:
    def __init__(self, so):
        L = self.lib = C.CDLL(so)
        for name, res, args in (("rstub_real", C.c_void_p, [C.c_void_p, C.c_long]), ("rstub_matrix", C.c_void_p, [C.c_void_p, C.c_int, C.c_int]),
                                ("rstub_int", C.c_void_p, [C.c_void_p, C.c_long]), ("rstub_lgl", C.c_void_p, [C.c_int]),
                                ("rstub_list", C.c_void_p, []), ("rstub_list_set", None, [C.c_void_p, C.c_char_p, C.c_void_p]),
                                ("rstub_type", C.c_int, [C.c_void_p]), ("rstub_length", C.c_long, [C.c_void_p]),
                                ("rstub_real_ptr", C.POINTER(C.c_double), [C.c_void_p]), ("rstub_int_ptr", C.POINTER(C.c_int), [C.c_void_p]),
                                ("rstub_text", C.c_char_p, [C.c_void_p]), ("rstub_get", C.c_void_p, [C.c_void_p, C.c_char_p]),
                                ("rstub_name", C.c_char_p, [C.c_void_p, C.c_long])):
            f = getattr(L, name)
            f.restype, f.argtypes = res, args

    # ---- Python value -> SEXP with R's types ----------------------------------------------------------------------------
    def to_sexp(self, v):
        L = self.lib
        if isinstance(v, dict):                                   # list(maxit = ..., eps_abs = ...)
            s = L.rstub_list()
            for k, x in v.items():
                L.rstub_list_set(s, k.encode(), self.to_sexp(x))
            return s
        if isinstance(v, (bool, np.bool_)):                       # logical(1)
            return L.rstub_lgl(int(v))
        if isinstance(v, (int, np.integer)):                      # integer(1)
            a = np.array([v], dtype=np.int32)
            return L.rstub_int(a.ctypes.data, 1)
        if isinstance(v, float):                                  # numeric(1)
            a = np.array([v], dtype=np.float64)
            return L.rstub_real(a.ctypes.data, 1)
        a = np.asarray(v)
        if a.ndim == 2:                                           # numeric matrix, column-major
            a = np.asfortranarray(a, dtype=np.float64)
            return L.rstub_matrix(a.ctypes.data, a.shape[0], a.shape[1])
        if a.dtype.kind in "iu":
            a = np.ascontiguousarray(a, dtype=np.int32)
            return L.rstub_int(a.ctypes.data, a.size)
        a = np.ascontiguousarray(a, dtype=np.float64)             # numeric vector (numeric(0) allowed)
        return L.rstub_real(a.ctypes.data, a.size)

    # ---- SEXP -> Python value ------------------------------------------------------------------------------------------
    def from_sexp(self, s):
        L = self.lib
        t = L.rstub_type(s)
        if t == ERROR:
            raise RError(L.rstub_text(s).decode())
        if t == NIL:
            return None
        if t == REAL:
            n = L.rstub_length(s)
            return np.array(L.rstub_real_ptr(s)[:n], dtype=np.float64)
        if t in (INT, LGL):
            n = L.rstub_length(s)
            return np.array(L.rstub_int_ptr(s)[:n], dtype=np.int32)
        items = {L.rstub_name(s, i).decode(): self.from_sexp(L.rstub_get(s, L.rstub_name(s, i))) for i in range(L.rstub_length(s))}
        if t == S4:
            items["class"] = L.rstub_text(s).decode()
        return items

    def dot_call(self, name, *args):
        """.Call(name, ...): every argument converted to the SEXP R would pass."""
        f = getattr(self.lib, name)
        f.restype, f.argtypes = C.c_void_p, [C.c_void_p] * len(args)
        return self.from_sexp(f(*[self.to_sexp(a) for a in args]))


def dgc_to_dense(m):
    """Matrix::dgCMatrix slots (Dim, p, i, x) -> dense array."""
    assert m["class"] == "dgCMatrix"
    nrow, ncol = int(m["Dim"][0]), int(m["Dim"][1])
    out = np.zeros((nrow, ncol))
    for k in range(ncol):
        sl = slice(int(m["p"][k]), int(m["p"][k + 1]))
        out[m["i"][sl], k] = m["x"][sl]
    return out


# ---- the fit() methods of the reference's R front end, restated call for call (argument order and types of
# R/30_admm_lasso.R:131-160, R/40_admm_enet.R:47-64, R/20_admm_lad.R:55-67, R/10_admm_bp.R:98-118) --------------------------
def r_opts(maxit=10000, eps_abs=1e-5, eps_rel=1e-5, rho=None):
    return {"maxit": int(maxit), "eps_abs": float(eps_abs), "eps_rel": float(eps_rel), "rho": -1.0 if rho is None else float(rho)}


def r_admm_lasso_fit(glue, x, y, lam=(), nlambda=100, lambda_min_ratio=None, standardize=True, intercept=True, nthread=1, **opts):
    lmr = float(lambda_min_ratio) if lambda_min_ratio is not None else (0.01 if x.shape[0] < x.shape[1] else 0.0001)
    lam = np.sort(np.asarray(lam, dtype=np.float64))[::-1]
    if nthread <= 1:
        return glue.dot_call("admm_lasso", x, y, lam, int(nlambda), lmr, bool(standardize), bool(intercept), r_opts(**opts))
    return glue.dot_call("admm_parlasso", x, y, lam, int(nlambda), lmr, bool(standardize), bool(intercept), int(nthread), r_opts(**opts))


def r_admm_enet_fit(glue, x, y, lam=(), alpha=0.5, nlambda=100, lambda_min_ratio=None, standardize=True, intercept=True, **opts):
    lmr = float(lambda_min_ratio) if lambda_min_ratio is not None else (0.01 if x.shape[0] < x.shape[1] else 0.0001)
    lam = np.sort(np.asarray(lam, dtype=np.float64))[::-1]
    return glue.dot_call("admm_enet", x, y, lam, int(nlambda), lmr, bool(standardize), bool(intercept), float(alpha), r_opts(**opts))


def r_admm_lad_fit(glue, x, y, intercept=True, maxit=10000, eps_abs=1e-4, eps_rel=1e-4, rho=1.0):
    return glue.dot_call("admm_lad", x, y, bool(intercept), r_opts(maxit, eps_abs, eps_rel, rho))


def r_admm_bp_fit(glue, x, y, maxit=10000, eps_abs=1e-4, eps_rel=1e-4, rho=1.0):
    return glue.dot_call("admm_bp", x, y, r_opts(maxit, eps_abs, eps_rel, rho))
