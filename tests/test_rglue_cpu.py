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
