"""CPU-side checks of the drop-in boundary: the shared library builds and loads, exports every
symbol include/b200admm.h declares, refuses to compute without a GPU (no CPU fallback), and the
Python mirror of the R front-end keeps the reference's defaults and stop() conditions."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def K():
    from admm_b200 import build, _capi
    build.build()
    return _capi


def header_functions():
    txt = open(os.path.join(ROOT, "include", "b200admm.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200admm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(K):
    L = K.lib()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert sorted(names) == sorted(K.EXPORTS)
    assert L.b200admm_version() == 100


def test_struct_layouts_match_header(K):
    # sizes a C compiler gives the header's structs on LP64
    assert C.sizeof(K.Data) == 40
    assert C.sizeof(K.Opts) == 32
    assert C.sizeof(K.Timing) == 64
    assert C.sizeof(K.Path) == 8 + 8 + 5 * 8 + 3 * 8 + 64
    assert C.sizeof(K.Dense) == 8 + 8 + 8 + 8 + 64


def test_no_cpu_fallback(K):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import admm_b200
    x = np.random.default_rng(0).normal(size=(50, 5))
    y = x[:, 0] + 0.1
    with pytest.raises(admm_b200.B200AdmmError) as e:
        admm_b200.admm_lasso(x, y).penalty([0.1]).fit()
    assert e.value.code == -2                      # B200ADMM_ENODEVICE
    with pytest.raises(admm_b200.B200AdmmError):
        admm_b200.device_info()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "admm_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


def test_front_end_defaults_and_validation():
    import admm_b200 as A
    x = np.zeros((30, 12)); y = np.zeros(30)
    m = A.admm_lasso(x, y)
    assert (m.nlambda, m.lambda_min_ratio, m.nthread, m.maxit, m.eps_abs, m.eps_rel, m.rho) == \
           (100, 1e-4, 1, 10000, 1e-5, 1e-5, -1.0)
    assert A.admm_lasso(np.zeros((5, 12)), np.zeros(5)).lambda_min_ratio == 0.01
    # a default goes to the library as 0 (resolved there from the GLOBAL row count of a row-sharded run), a user's value as is
    assert m._lmr_arg() == 0.0 and m.penalty(nlambda=5)._lmr_arg() == 0.0
    assert m.penalty(lambda_min_ratio=0.5)._lmr_arg() == 0.5
    assert m.penalty([0.1, 0.5, 0.2]).lambda_.tolist() == [0.5, 0.2, 0.1]      # sorted decreasingly
    assert m.penalty() is m and m.opts() is m and m.parallel(2) is m
    for bad in (lambda: m.penalty([-1.0]), lambda: m.penalty(nlambda=0), lambda: m.penalty(lambda_min_ratio=1.0),
                lambda: m.opts(maxit=0), lambda: m.opts(eps_abs=-1), lambda: m.opts(rho=0.0),
                lambda: m.parallel(3),                                         # nthread >= ncol/5
                lambda: A.admm_lasso(x, np.zeros(29)),
                lambda: A.admm_enet(x, y).penalty(alpha=1.5),
                lambda: A.admm_lad(np.zeros((5, 12)), np.zeros(5)),            # needs n > p
                lambda: A.admm_bp(x, y),                                       # needs p > n
                lambda: A.admm_bp(np.zeros((5, 12)), np.zeros(5)).opts(rho=0)):
        with pytest.raises(ValueError):
            bad()
    lad = A.admm_lad(x, y)
    assert (lad.eps_abs, lad.eps_rel, lad.rho, lad.maxit) == (1e-4, 1e-4, 1.0, 10000)
    bp = A.admm_bp(np.zeros((5, 12)), np.zeros(5))
    assert (bp.eps_abs, bp.rho) == (1e-4, 1.0)
    with pytest.raises(RuntimeError):
        bp.parallel(2).fit()                      # admm_parbp is not part of the reference build either


def test_host_lanczos_matches_oracle_on_cpu():
    """The product's host Lanczos driver (coarse_eig.hpp) against the oracle's, same operator."""
    import subprocess, tempfile, textwrap
    from oracle import pyoracle as O
    rng = np.random.default_rng(7)
    x = rng.normal(size=(300, 40)).astype(np.float32)
    G = np.asfortranarray((x.T @ x).astype(np.float32))
    ev_o, info = O.coarse_eig_f32(G)
    src = textwrap.dedent(r'''
        #include "coarse_eig.hpp"
        #include <cstdio>
        int main(int argc, char** argv) {
            int n; if (fread(&n, 4, 1, stdin) != 1) return 1;
            std::vector<float> S((size_t)n * n);
            if (fread(S.data(), 4, S.size(), stdin) != S.size()) return 1;
            b200::eig::LanczosInfo li;
            auto op = [&](const float* v, float* w) {       // same arithmetic order as the oracle's symv
                for (int i = 0; i < n; i++) w[i] = 0.f;
                for (int j = 0; j < n; j++) { float s = 0.f;
                    for (int i = 0; i < n; i++) s += S[(size_t)j * n + i] * v[i]; w[j] = s; }
            };
            float ev = b200::eig::coarse_largest_eigenvalue<float>(op, n, &li);
            printf("%.9g %d %d %d\n", ev, li.nmatvec, li.nrestart, li.converged);
            return 0;
        }''')
    with tempfile.TemporaryDirectory() as td:
        cpp = os.path.join(td, "t.cpp")
        open(cpp, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-I", os.path.join(ROOT, "admm_b200", "csrc"), cpp, "-o", exe])
        Gs = np.tril(G) + np.tril(G, -1).T
        inp = np.int32(40).tobytes() + np.asfortranarray(Gs.astype(np.float32)).tobytes(order="F")
        out = subprocess.run([exe], input=inp, capture_output=True, check=True).stdout.decode().split()
    ev_p = float(out[0])
    assert abs(ev_p - ev_o) < 2e-5 * ev_o          # same algorithm; the mat-vec sums in another order
    assert int(out[1]) == info["nmatvec"] and int(out[3]) == info["converged"]


def test_header_is_plain_c_and_a_c_program_links(tmp_path, K):
    """The drop-in boundary is a C ABI: the header must compile as C99 and as C++, and a plain C program must
    link against the shared library and resolve every entry point (it only asks for the version and, without a
    GPU, gets B200ADMM_ENODEVICE from a solver call -- no compute on CPU)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no C compiler")
    hdr = os.path.join(ROOT, "include", "b200admm.h")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    subprocess.run([shutil.which("g++") or gcc, "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "b200admm.h"
int main(void) {
    double x[6] = {1, 2, 3, 4, 5, 7}, y[3] = {1, 2, 3}, lam[1] = {0.1};
    b200admm_data d; b200admm_opts o; b200admm_path P;
    memset(&d, 0, sizeof d); memset(&P, 0, sizeof P);
    d.n = 3; d.p = 2; d.dtype = B200ADMM_F64_HOST; d.x = x; d.y = y;
    o.maxit = 10; o.eps_abs = 1e-5; o.eps_rel = 1e-5; o.rho = -1.0;
    int rc = b200admm_lasso(&d, lam, 1, 1, 1e-4, 1, 1, &o, &P);
    printf("%d %d %s\n", b200admm_version(), rc, rc ? b200admm_last_error() : "ok");
    if (rc == 0) b200admm_free_path(&P);
    return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(K.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libb200admm.so", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split(None, 2)
    assert out[0] == "100"
    import torch
    if not torch.cuda.is_available():
        assert int(out[1]) == -2, out                     # B200ADMM_ENODEVICE: no CPU fallback behind the ABI
