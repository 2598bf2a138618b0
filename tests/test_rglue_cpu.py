"""The Rcpp glue a maintainer of the reference adds (r-pkg/src/b200_glue.cpp) type-checks against include/b200admm.h
and exports the five `.Call` entry points of the reference with their argument counts.  R and Rcpp are not in this
image: <Rcpp.h> is a minimal stand-in (tests/stubs/Rcpp.h) that declares only what the glue uses."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_R = "/root/reference/R"


def test_glue_compiles_and_exports_the_five_entry_points(tmp_path):
    obj = str(tmp_path / "b200_glue.o")
    r = subprocess.run(["g++", "-std=c++14", "-Wall", "-Werror", "-c", os.path.join(ROOT, "r-pkg", "src", "b200_glue.cpp"),
                        "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"), "-o", obj],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    nm = subprocess.run(["nm", "-g", "--defined-only", obj], capture_output=True, text=True).stdout
    exported = sorted(line.split()[-1] for line in nm.splitlines() if " T " in line)
    assert exported == ["admm_bp", "admm_enet", "admm_lad", "admm_lasso", "admm_parlasso"]
    undefined = subprocess.run(["nm", "-u", obj], capture_output=True, text=True).stdout
    for sym in ("b200admm_lasso", "b200admm_enet", "b200admm_parlasso", "b200admm_lad", "b200admm_bp", "b200admm_free_path",
                "b200admm_free_dense", "b200admm_last_error"):
        assert re.search(r"\bU %s\b" % sym, undefined), sym


def test_glue_in_integration_md_is_the_file_in_r_pkg():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    glue = open(os.path.join(ROOT, "r-pkg", "src", "b200_glue.cpp")).read()
    assert glue.strip() in text


def test_argument_counts_match_the_reference_r_files():
    """.Call("admm_lasso", x, y, lambda, nlambda, lambda_min_ratio, standardize, intercept, opts, PACKAGE = "ADMM") etc."""
    glue = open(os.path.join(ROOT, "r-pkg", "src", "b200_glue.cpp")).read()
    nargs = {m.group(1): m.group(2).count("SEXP") for m in re.finditer(r"RcppExport SEXP (\w+)\(([^)]*)\)", glue)}
    assert nargs == {"admm_lasso": 8, "admm_enet": 9, "admm_parlasso": 9, "admm_lad": 4, "admm_bp": 3}
    if not os.path.isdir(REF_R):
        return                                             # the reference tree exists in the build container only
    calls = {}
    for fn in os.listdir(REF_R):
        src = open(os.path.join(REF_R, fn)).read()
        for m in re.finditer(r'\.Call\("(\w+)",(.*?)PACKAGE\s*=\s*"ADMM"\)', src, re.S):
            args = re.sub(r"list\([^)]*\)", "opts", m.group(2))                  # the opts list is ONE argument
            calls[m.group(1)] = len([a for a in args.split(",") if a.strip()])
    for name, n in nargs.items():
        assert calls.get(name) == n, (name, calls.get(name), n)


# ---- the glue EXECUTED: functional <Rcpp.h> stand-in + a recording fake of the library (tests/stubs/) -------------------------
import ctypes as C                                          # noqa: E402
import sys                                                  # noqa: E402

import numpy as np                                          # noqa: E402
import pytest                                               # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import rstub                                                # noqa: E402


class FakeRecord(C.Structure):
    _fields_ = [("entry", C.c_char * 32), ("n", C.c_longlong), ("p", C.c_longlong), ("dtype", C.c_int), ("x", C.c_void_p), ("y", C.c_void_p),
                ("lambda_", C.c_double * 8), ("nlambda_given", C.c_int), ("nlambda", C.c_int), ("lmin_ratio", C.c_double),
                ("standardize", C.c_int), ("intercept", C.c_int), ("alpha", C.c_double), ("nthread", C.c_int),
                ("maxit", C.c_int), ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("rho", C.c_double),
                ("frees_path", C.c_int), ("frees_dense", C.c_int), ("fail", C.c_int)]


@pytest.fixture(scope="module")
def fake(tmp_path_factory):
    g = rstub.build(tmp_path_factory.mktemp("rglue"), fake=True)
    g.lib.fake_record.restype = C.POINTER(FakeRecord)
    return g


def rec(g):
    return g.lib.fake_record().contents


def test_lasso_call_marshals_the_r_arguments_in_the_reference_order(fake):
    """R/30_admm_lasso.R:140-147 -> src/Lasso.cpp:32-35: x, y, lambda, nlambda, lambda_min_ratio, standardize, intercept, opts."""
    fake.lib.fake_reset()
    x = np.arange(12.0).reshape(4, 3)
    y = np.arange(4.0)
    res = rstub.r_admm_lasso_fit(fake, x, y, lam=[0.1, 0.5, 0.3], nlambda=7, standardize=False, intercept=True, maxit=123, eps_abs=1e-3, eps_rel=2e-3, rho=4.5)
    r = rec(fake)
    assert r.entry == b"lasso" and (r.n, r.p, r.dtype) == (4, 3, 0)                        # B200ADMM_F64_HOST = R's REALSXP
    assert (r.nlambda_given, r.nlambda, r.standardize, r.intercept) == (3, 7, 0, 1) and list(r.lambda_[:3]) == [0.5, 0.3, 0.1]
    assert (r.maxit, r.eps_abs, r.eps_rel, r.rho, r.lmin_ratio) == (123, 1e-3, 2e-3, 4.5, 0.0001)
    xs = np.ctypeslib.as_array(C.cast(r.x, C.POINTER(C.c_double)), (12,))
    assert np.array_equal(xs, x.ravel(order="F")) and r.frees_path == 1                    # column-major, result released once
    # List(lambda, beta = dgCMatrix, niter): src/Lasso.cpp:131-135
    assert list(res) == ["lambda", "beta", "niter"]
    assert np.array_equal(res["lambda"], [1.0, 0.5, 1.0 / 3]) and np.array_equal(res["niter"], [10, 11, 12]) and res["niter"].dtype == np.int32
    b = rstub.dgc_to_dense(res["beta"])
    assert b.shape == (4, 3) and np.array_equal(b[0], [0.5, 1.5, 2.5]) and np.array_equal(np.diag(b[1:]), [-1.0, -2.0, -3.0])
    assert res["beta"]["p"].dtype == np.int32 and np.array_equal(res["beta"]["p"], [0, 2, 4, 6])


def test_lasso_without_lambdas_passes_numeric0_and_the_rho_default(fake):
    fake.lib.fake_reset()
    x = np.zeros((3, 5))
    res = rstub.r_admm_lasso_fit(fake, x, np.zeros(3), nlambda=4)
    r = rec(fake)
    assert (r.nlambda_given, r.nlambda, r.rho, r.lmin_ratio) == (0, 4, -1.0, 0.01)        # n < p default; rho = NULL -> -1
    assert rstub.dgc_to_dense(res["beta"]).shape == (6, 4)


def test_enet_parlasso_lad_bp_calls(fake):
    x = np.arange(20.0).reshape(5, 4)
    y = np.ones(5)
    fake.lib.fake_reset()
    rstub.r_admm_enet_fit(fake, x, y, lam=[0.2], alpha=0.3)
    assert rec(fake).entry == b"enet" and rec(fake).alpha == 0.3 and rec(fake).nlambda_given == 1
    fake.lib.fake_reset()
    rstub.r_admm_lasso_fit(fake, x, y, lam=[0.2], nthread=2)
    assert rec(fake).entry == b"parlasso" and rec(fake).nthread == 2
    fake.lib.fake_reset()
    res = rstub.r_admm_lad_fit(fake, x, y, intercept=False)
    assert rec(fake).entry == b"lad" and rec(fake).intercept == 0 and (rec(fake).eps_abs, rec(fake).rho) == (1e-4, 1.0)
    assert list(res) == ["beta", "niter"] and np.array_equal(res["beta"], 0.25 * np.arange(5)) and int(res["niter"][0]) == 77   # src/LAD.cpp:44-45
    assert rec(fake).frees_dense == 1
    fake.lib.fake_reset()
    res = rstub.r_admm_bp_fit(fake, x, y)
    assert rec(fake).entry == b"bp" and list(res) == ["beta", "niter"]                     # src/BP.cpp:38-43: no lambda, scalar niter
    assert rstub.dgc_to_dense(res["beta"]).shape == (4, 1) and res["niter"].shape == (1,) and int(res["niter"][0]) == 10


def test_a_library_error_becomes_an_r_error_with_the_library_message(fake):
    fake.lib.fake_reset()
    rec(fake).fail = -2
    with pytest.raises(rstub.RError, match="requested failure"):
        rstub.r_admm_lasso_fit(fake, np.zeros((3, 2)), np.zeros(3))
    assert rec(fake).frees_path == 0


def test_bad_argument_types_raise_r_errors_before_the_library_is_called(fake):
    fake.lib.fake_reset()
    with pytest.raises(rstub.RError):
        fake.dot_call("admm_lasso", np.zeros(6), np.zeros(3), np.zeros(0), 5, 0.01, True, True, rstub.r_opts())    # x is not a matrix
    with pytest.raises(rstub.RError, match="maxit"):
        fake.dot_call("admm_lasso", np.zeros((3, 2)), np.zeros(3), np.zeros(0), 5, 0.01, True, True, {"eps_abs": 1e-5})
    assert rec(fake).entry == b""


def test_glue_against_the_real_library_fails_loudly_without_a_gpu(tmp_path):
    """No CPU fallback behind the R entry points either: on a box without a CUDA device the .Call raises the library's error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: tests/test_gpu_rglue.py runs the glue for real")
    g = rstub.build(tmp_path, fake=False)
    with pytest.raises(rstub.RError, match="CUDA"):
        rstub.r_admm_lasso_fit(g, np.random.default_rng(0).normal(size=(30, 4)), np.zeros(30))
