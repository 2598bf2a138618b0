"""GPU unit tests of the individual CUDA kernels behind the C ABI (b200admm_k_* entry points),
each against a float64 NumPy statement of the same operation or against the CPU oracle.
Floating-point tolerances are stated per test; integer/index outputs must match exactly."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from admm_b200 import _capi
    _capi.device_info()
    return _capi


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def colmajor_dev(torch, a):
    """numpy (n, p) -> CUDA tensor holding the column-major image (a contiguous (p, n) tensor)."""
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()


def back(t):
    return t.cpu().numpy().T


@pytest.mark.parametrize("n,p", [(1000, 64), (777, 33), (4096, 5), (65, 130)])
@pytest.mark.parametrize("flags", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_standardize_matches_oracle(K, torch, n, p, flags):
    from oracle import pyoracle as O
    rng = np.random.default_rng(n * p)
    x = rng.normal(1.2, 2.0, size=(n, p)).astype(np.float32)
    y = (rng.normal(size=n) + 5).astype(np.float32)
    xd, yd = colmajor_dev(torch, x), torch.from_numpy(y.copy()).cuda()
    out = torch.empty_like(xd)
    meanx = np.zeros(p, np.float32); scalex = np.ones(p, np.float32); ys = np.zeros(2, np.float32)
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_k_standardize_f32(xd.data_ptr(), out.data_ptr(), yd.data_ptr(), n, p, flags[0], flags[1],
                                               meanx.ctypes.data, scalex.ctypes.data, ys.ctypes.data))
    xo = np.asfortranarray(x.copy()); yo = y.copy()
    st = O.standardize_f32(xo, yo, standardize=bool(flags[0]), intercept=bool(flags[1]))
    # float32 data, sums taken in a different order: a few ulp
    assert np.allclose(back(out), xo, rtol=2e-5, atol=2e-6)
    assert np.allclose(yd.cpu().numpy(), yo, rtol=2e-5, atol=2e-6)
    if flags[1]:
        assert np.allclose(meanx, st["meanX"], rtol=1e-5, atol=1e-6)
        assert abs(ys[0] - st["meanY"]) < 1e-5 * max(1, abs(st["meanY"]))
    if flags[0]:
        assert np.allclose(scalex, st["scaleX"], rtol=1e-5)
    if flags[0] or flags[1]:
        assert abs(ys[1] - st["scaleY"]) < 1e-5 * st["scaleY"]
    assert np.array_equal(back(xd), x)          # input untouched (out-of-place)


@pytest.mark.parametrize("n,p", [(5000, 128), (1234, 77), (300, 300), (20000, 260)])
def test_gram_cuda_cores(K, torch, n, p):
    rng = np.random.default_rng(n + p)
    x = rng.normal(size=(n, p)).astype(np.float32)
    xd = colmajor_dev(torch, x)
    g = torch.zeros(p, p, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_k_gram_f32(xd.data_ptr(), n, p, g.data_ptr(), 0))
    ref = x.astype(np.float64).T @ x.astype(np.float64)
    got = g.cpu().numpy()
    assert np.array_equal(got, got.T)                                   # exactly symmetric (mirrored store)
    # float32 accumulation of n terms: error ~ sqrt(n) * eps * |x_i||x_j|
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    assert (np.abs(got - ref) / scale).max() < 3e-6 * max(1.0, np.sqrt(n / 1000.0))


def test_gemv_t(K, torch):
    rng = np.random.default_rng(4)
    # (33001 x 9600: long columns AND more columns than CTA slots -- the shape class of the consensus solver's K_i^-1 product,
    #  cut into two row segments per column; odd length: scalar tail)
    for m, nc in [(100000, 7), (1000, 1000), (37, 5), (40001, 3), (33001, 9600)]:
        a = rng.normal(size=(m, nc)).astype(np.float32)
        v = rng.normal(size=m).astype(np.float32)
        ad, vd = colmajor_dev(torch, a), torch.from_numpy(v).cuda()
        out = torch.zeros(nc, device="cuda", dtype=torch.float32)
        torch.cuda.synchronize()
        K.check(K.lib().b200admm_k_gemv_t_f32(ad.data_ptr(), m, nc, vd.data_ptr(), out.data_ptr()))
        ref = a.astype(np.float64).T @ v.astype(np.float64)
        assert np.abs(out.cpu().numpy() - ref).max() < 2e-6 * np.sqrt(m) * 3


@pytest.mark.parametrize("p", [64, 128, 129, 500, 1000, 1300, 2100, 4352])
def test_cholesky_and_spd_inverse(K, torch, p):
    rng = np.random.default_rng(p)
    x = rng.normal(size=(4 * p, p))
    a = (x.T @ x + 5.0 * np.eye(p)).astype(np.float32)
    ad = torch.from_numpy(a.copy()).cuda()               # symmetric: row/column major coincide
    info = C.c_int(-1)
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_k_chol_f32(ad.data_ptr(), p, C.byref(info)))
    assert info.value == 0
    L = np.tril(back(ad)).astype(np.float64)
    a64 = a.astype(np.float64)
    assert np.abs(L @ L.T - a64).max() / np.abs(a64).max() < 5e-6      # backward error of a float32 factorisation
    # explicit inverse
    ad2 = torch.from_numpy(a.copy()).cuda()
    work = torch.empty(p, p, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_k_spd_inverse_f32(ad2.data_ptr(), p, work.data_ptr(), C.byref(info)))
    inv = ad2.cpu().numpy().astype(np.float64)
    assert np.array_equal(inv, inv.T)
    cond = np.linalg.cond(a64)
    assert np.abs(inv @ a64 - np.eye(p)).max() < 2e-6 * cond * 4


def test_cholesky_reports_indefinite(K, torch):
    p = 200
    a = np.eye(p, dtype=np.float32)
    a[150, 150] = -1.0
    ad = torch.from_numpy(a).cuda()
    info = C.c_int(0)
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_k_chol_f32(ad.data_ptr(), p, C.byref(info)))
    assert info.value == 151


def numpy_fused_zu(x, adj_y, old_z, adj_z, lam, rho, enet, alpha):
    """float32 restatement of ADMMLassoTall::next_z / next_residual + FADMMBase::update_y."""
    f = np.float32
    frho = f(rho)
    v = (x + adj_y / frho).astype(f)
    pen = float(f(lam)) / rho
    if not enet:
        vd = v.astype(np.float64)
        z = np.where(vd > pen, vd - pen, np.where(vd < -pen, vd + pen, 0.0)).astype(f)
    else:
        th = f(float(f(alpha)) * pen)
        den = f(1.0 + pen * (1.0 - float(f(alpha))))
        z = np.where(v > th, (v - th) / den, np.where(v < -th, (v + th) / den, f(0))).astype(f)
    r = (x - z).astype(f)
    y = (adj_y + frho * r).astype(f)
    d = lambda a: float(np.sum(a.astype(np.float64) ** 2))
    return z, y, [d(r), d(z - old_z), d(z - adj_z), d(x), d(z), d(y)]


@pytest.mark.parametrize("enet", [0, 1])
@pytest.mark.parametrize("length", [1 << 20, 1000003, 5])
def test_fused_zu_kernel(K, torch, enet, length):
    rng = np.random.default_rng(length + enet)
    x, ay, oz, az = (rng.normal(size=length).astype(np.float32) for _ in range(4))
    oz[rng.uniform(size=length) < 0.5] = 0
    dev = [torch.from_numpy(a).cuda() for a in (x, ay, oz, az)]
    z = torch.empty(length, device="cuda", dtype=torch.float32)
    y = torch.empty(length, device="cuda", dtype=torch.float32)
    sums = np.zeros(6)
    ms = C.c_float(0)
    lam, rho, alpha = 0.7, 1.3, 0.4
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_k_fused_zu_f32(*(t.data_ptr() for t in dev), z.data_ptr(), y.data_ptr(), length,
                                            lam, rho, enet, alpha, sums.ctypes.data, C.byref(ms), 1))
    zr, yr, sr = numpy_fused_zu(x, ay, oz, az, lam, rho, enet, alpha)
    assert np.array_equal(z.cpu().numpy(), zr)        # element-wise part is bit-exact
    assert np.array_equal(y.cpu().numpy(), yr)
    assert np.allclose(sums, sr, rtol=2e-5)           # float32 partial sums, different order


def test_synth_design_statistics_and_sharding(K, torch):
    n, p = 40000, 64
    x = torch.empty(p, n, device="cuda", dtype=torch.float32)
    y = torch.empty(n, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_synth_f32(x.data_ptr(), y.data_ptr(), n, p, 0, 123, 0.0, 2.0, 10, 1.0))
    xa = x.cpu().numpy()
    assert abs(xa.mean()) < 0.02 and abs(xa.std() - 2.0) < 0.02
    assert abs(np.corrcoef(xa[0], xa[1])[0, 1]) < 0.03
    # a row block generated on its own is the same rows of the same global matrix
    r0, nr = 10001, 5003
    xs = torch.empty(p, nr, device="cuda", dtype=torch.float32)
    ys = torch.empty(nr, device="cuda", dtype=torch.float32)
    K.check(K.lib().b200admm_synth_f32(xs.data_ptr(), ys.data_ptr(), nr, p, r0, 123, 0.0, 2.0, 10, 1.0))
    assert np.array_equal(xs.cpu().numpy(), xa[:, r0:r0 + nr])
    assert np.array_equal(ys.cpu().numpy(), y.cpu().numpy()[r0:r0 + nr])


@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("n,p", [(4096, 256), (20000, 300), (1000, 8), (50000, 129), (16, 40), (2048 + 16, 512)])
def test_gram_tensor_cores(K, torch, n, p, mode):
    """tcgen05 split-product Gram (3xTF32, 3xFP16) against float64: at least as accurate as the float32 CUDA-core kernel."""
    rng = np.random.default_rng(n + p + mode)
    x = (rng.normal(0.3, 2.0, size=(n, p))).astype(np.float32)
    xd = colmajor_dev(torch, x)
    g = torch.full((p, p), 7.0, device="cuda", dtype=torch.float32)     # the entry point zeroes it
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_k_gram_f32(xd.data_ptr(), n, p, g.data_ptr(), mode))
    ref = x.astype(np.float64).T @ x.astype(np.float64)
    got = g.cpu().numpy()
    assert np.array_equal(got, got.T)
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    assert (np.abs(got - ref) / scale).max() < 3e-6 * max(1.0, np.sqrt(n / 1000.0))


def test_gram_fp16_split_flags_out_of_range_values(K, torch):
    """The fp16 hi / lo split is only valid for unit-scale data: a value beyond the fp16 range must be
    reported, not silently turned into inf (the solvers use this split only behind DataStd flag 3, where
    |x| <= sqrt(n) holds by construction)."""
    from admm_b200 import B200AdmmError
    n, p = 4096, 256
    x = torch.randn((p, n), device="cuda", dtype=torch.float32)
    x[3, 100] = 1.0e5
    g = torch.zeros((p, p), device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    with pytest.raises(B200AdmmError):
        K.check(K.lib().b200admm_k_gram_f32(x.data_ptr(), n, p, g.data_ptr(), 3))
    # the flag is cleared by the failed call: the next clean call succeeds, and TF32 takes the same data
    x[3, 100] = 1.0
    K.check(K.lib().b200admm_k_gram_f32(x.data_ptr(), n, p, g.data_ptr(), 3))
    x[3, 100] = 1.0e5
    K.check(K.lib().b200admm_k_gram_f32(x.data_ptr(), n, p, g.data_ptr(), 2))
    ref = x.double() @ x.double().t()
    assert ((g.double() - ref).abs() / torch.sqrt(torch.outer(torch.diag(ref), torch.diag(ref)))).max() < 3e-6


def test_gram_fp16_split_small_magnitudes(K, torch):
    """Columns of very different magnitude inside the unit-scale envelope: the lo half turns subnormal
    (absolute precision 2^-25) and the result must still match float64 at float32 level on the unit scale."""
    n, p = 8192, 384
    g0 = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((p, n), device="cuda", dtype=torch.float32, generator=g0)
    x[::3] *= 1e-3                                                       # every third column is tiny
    g = torch.zeros((p, p), device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_k_gram_f32(x.data_ptr(), n, p, g.data_ptr(), 3))
    ref = x.double() @ x.double().t()
    assert (g.double() - ref).abs().max() < 3e-6 * n                     # absolute, relative to unit-scale entries ~ n


def test_gram_tensor_rejects_unaligned(K, torch):
    from admm_b200 import B200AdmmError
    x = torch.zeros((64, 1001), device="cuda", dtype=torch.float32)     # n = 1001 is not a multiple of 4
    g = torch.zeros((64, 64), device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    with pytest.raises(B200AdmmError):
        K.check(K.lib().b200admm_k_gram_f32(x.data_ptr(), 1001, 64, g.data_ptr(), 1))


def test_product_coarse_eig_matches_independent_numpy_restatement():
    """The product's Lanczos driver (coarse_eig.hpp, device mat-vecs) against the NumPy restatement of Spectra
    (tests/spectra_numpy.py) -- not against the oracle, whose Lanczos is a textual sibling of the product's.
    25 matrices, 0-3 implicit restarts; 5e-6 relative, same operator-application and restart counts."""
    import ctypes as C
    import torch
    import lanczos_cases
    import spectra_numpy as SN
    from admm_b200 import _capi as K
    L = K.lib()
    restarts = set()
    worst = 0.0
    for name, S in lanczos_cases.cases():
        Sd = torch.from_numpy(np.ascontiguousarray(S)).cuda()
        ev = C.c_float(0)
        info = np.zeros(3, dtype=np.int32)
        torch.cuda.synchronize()
        K.check(L.b200admm_k_coarse_eig_f32(Sd.data_ptr(), S.shape[0], C.byref(ev), info.ctypes.data))
        ev_n, nmat, nrs, conv = SN.coarse_largest_eigenvalue(S)
        assert info[2] == conv == 1, name
        assert info[0] == nmat and info[1] == nrs, (name, info, nmat, nrs)
        worst = max(worst, abs(ev.value / ev_n - 1.0))
        assert abs(ev.value / ev_n - 1.0) < 5e-6, (name, ev.value, ev_n)
        restarts.add(nrs)
    print("\n[parity] product Lanczos vs independent NumPy restatement: worst relative difference %.2e over 25 matrices" % worst)
    assert {0, 1, 2, 3} <= restarts


@pytest.mark.parametrize("M,N,Kd,tile_mode,klo,khi,epi", [
    (700, 900, 300, 0, 0, 0, 0),        # rectangle, ragged everywhere, store
    (512, 1024, 128, 0, 0, 0, 1),       # rank-128 update: C -= A'B
    (1000, 1000, 1030, 2, 0, 0, 1),     # tiles on and above the diagonal, subtract (Cholesky trailing update)
    (1100, 600, 600, 0, 2, 0, 0),       # B lower triangular: K from block J on
    (900, 520, 900, 0, 0, 1, 2),        # A upper triangular: K up to block I; C = -A'B
    (1300, 1300, 1300, 1, 3, 0, 0),     # Z'Z with Z lower triangular: tiles below the diagonal, K from max(I, J)
])
def test_tensor_core_tn_product(K, torch, M, N, Kd, tile_mode, klo, khi, epi):
    """tn_pair_kernel (3xTF32, CTA pairs) against float64: every tile set, K-range rule and epilogue the blocked
    factorisation uses.  Bound: 4e-6 * sqrt(K) * max|A| max|B| (fp32-accurate accumulation)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + Kd)
    lda, ldb, ldc = (Kd + 3) // 4 * 4 + 4, (Kd + 3) // 4 * 4, (M + 3) // 4 * 4
    A = torch.zeros((M, lda), device="cuda")            # column-major K x M: row i of this tensor = column i of A
    B = torch.zeros((N, ldb), device="cuda")
    A[:, :Kd] = torch.randn((M, Kd), device="cuda", generator=g)
    B[:, :Kd] = torch.randn((N, Kd), device="cuda", generator=g)
    ki = torch.arange(Kd, device="cuda")
    if klo == 2:                                        # B(k, j) = 0 for k < j
        B[:, :Kd] *= (ki[None, :] >= torch.arange(N, device="cuda")[:, None]).float()
    if khi == 1:                                        # A(k, i) = 0 for k > i
        A[:, :Kd] *= (ki[None, :] <= torch.arange(M, device="cuda")[:, None]).float()
    if klo == 3:                                        # both lower triangular (one matrix Z)
        A[:, :Kd] *= (ki[None, :] >= torch.arange(M, device="cuda")[:, None]).float()
        B = A
        ldb = lda
    C0 = torch.randn((N, ldc), device="cuda", generator=g)
    Cd = C0.clone()
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_k_gemm_tn_f32(A.data_ptr(), lda, B.data_ptr(), ldb, M, N, Kd, Cd.data_ptr(), ldc, tile_mode, klo, khi, epi))
    prod = (B[:, :Kd].double() @ A[:, :Kd].double().t())            # (N, M): entry (j, i) = C(i, j)
    want = {0: prod, 1: C0[:, :M].double() - prod, 2: -prod}[epi]
    got = Cd[:, :M].double()
    ii = torch.arange(M, device="cuda")[None, :] // 256
    jj = torch.arange(N, device="cuda")[:, None] // 256
    mask = {0: torch.ones_like(prod, dtype=torch.bool), 1: jj <= ii, 2: jj >= ii}[tile_mode]
    err = ((got - want).abs() * mask).max().item()
    assert err < 4e-6 * np.sqrt(Kd) * 16, err
    # tiles outside the requested set and the padding rows of C are untouched
    assert torch.equal(Cd[:, M:], C0[:, M:])
    if tile_mode != 0:
        assert torch.equal(Cd[:, :M][~mask], C0[:, :M][~mask])


@pytest.mark.parametrize("ta,tb,M,N,Kd,mode", [
    (1, 0, 1700, 1700, 3000, 3),        # X'X lower + mirror (LAD Gram): tensor-core tiles (14 x 14 >= 148)
    (0, 1, 1664, 1664, 2111, 3),        # A A' lower + mirror (BP Gram)
    (0, 0, 2000, 1900, 2000, 4),        # op(A) = A lower triangular (M = L^-1 A panels)
    (0, 1, 1800, 1700, 1700, 32),       # op(B) = B', B lower triangular
    (1, 0, 1900, 1800, 1900, 8),        # op(A) = A', A lower triangular
    (0, 0, 1800, 1700, 1800, 16),       # op(B) = B lower triangular
    (1, 1, 1601, 1555, 1033, 0),        # plain, ragged edges
    (0, 0, 300, 200, 150, 0),           # small: CUDA-core tiles
])
def test_gemm_f64_all_modes(K, torch, ta, tb, M, N, Kd, mode):
    """float64 product behind the LAD / BP setup (tensor-core mma.sync.f64 tiles when the grid fills the chip) against
    NumPy float64, every transposition / triangular-operand / lower-mirror mode; alpha, beta != trivial."""
    rng = np.random.default_rng(M + N + Kd + mode)
    a = rng.normal(size=(Kd, M) if ta else (M, Kd))
    b = rng.normal(size=(N, Kd) if tb else (Kd, N))
    if mode & 1:
        b = a.copy()                    # lower / mirror stores are for symmetric products (X'X, A A')
    opa = a.T if ta else a
    opb = b.T if tb else b
    if mode & 4:
        a = np.tril(a); opa = a
    if mode & 8:
        a = np.tril(a); opa = a.T
    if mode & 16:
        b = np.tril(b); opb = b
    if mode & 32:
        b = np.tril(b); opb = b.T
    c0 = rng.normal(size=(M, N))
    alpha, beta = 0.75, (0.0 if mode & 3 else -0.5)
    ref = alpha * (opa @ opb) + beta * c0
    # triangular operands: the kernel must not read the structurally-zero half -> poison it
    a_dev, b_dev = a.copy(), b.copy()
    if mode & (4 | 8):
        a_dev[np.triu_indices_from(a_dev, 1)] = np.nan
    if mode & (16 | 32):
        b_dev[np.triu_indices_from(b_dev, 1)] = np.nan
    ad = torch.from_numpy(np.ascontiguousarray(a_dev.T)).cuda()       # column-major on the device
    bd = torch.from_numpy(np.ascontiguousarray(b_dev.T)).cuda()
    cd = torch.from_numpy(np.ascontiguousarray(c0.T)).cuda()
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_k_gemm_f64(ta, tb, M, N, Kd, alpha, ad.data_ptr(), a.shape[0], bd.data_ptr(), b.shape[0], beta,
                                        cd.data_ptr(), M, mode, None, 1))
    got = cd.cpu().numpy().T
    if mode & 1:                        # lower (+ mirror): square C
        if mode & 2:
            assert np.array_equal(got, got.T)
            assert np.abs(got - ref).max() < 1e-11 * np.abs(ref).max()
        else:
            il = np.tril_indices(M)
            assert np.abs(got[il] - ref[il]).max() < 1e-11 * np.abs(ref).max()
    else:
        assert np.abs(got - ref).max() < 1e-11 * np.abs(ref).max()
