"""Host-side scheduling logic of the CUDA library, exercised on CPU through the C ABI's host-only planners
(no device call): the work distribution of the fp16 CTA-pair Gram kernel (full rounds of whole tiles, the last
round cut along K into canonical slices) and the column panels of the pipelined host ingest."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L():
    from admm_b200 import build, _capi
    build.build()
    return _capi.lib()


def plan(L, ntiles, npairs, nk):
    cover = np.zeros(ntiles * nk, dtype=np.int32)
    per_pair = np.zeros(npairs, dtype=np.int64)
    ns, st = C.c_int(0), C.c_int(0)
    rc = L.b200admm_k_gram_plan(ntiles, npairs, nk, cover.ctypes.data, per_pair.ctypes.data, C.byref(ns), C.byref(st))
    assert rc == 0
    return cover.reshape(ntiles, nk), per_pair, ns.value, st.value


@pytest.mark.parametrize("ntiles,npairs,nk", [
    (820, 74, 3125),      # the headline shape (p = 1e4: 40 row blocks), shortened K
    (136, 74, 625), (153, 74, 625), (210, 74, 100),
    (1, 74, 128), (3, 74, 32), (40, 74, 1000), (74, 74, 7), (75, 74, 5), (820, 74, 1), (6, 74, 3), (37, 74, 97),
    (5, 2, 50), (7, 1, 9),
])
def test_gram_plan_covers_every_stage_exactly_once(L, ntiles, npairs, nk):
    cover, per_pair, nslices, split_tiles = plan(L, ntiles, npairs, nk)
    assert (cover == 1).all(), "a (tile, stage) is computed %d times" % int(cover.max() if cover.max() != 1 else cover.min())
    assert int(per_pair.sum()) == ntiles * nk
    nch = (nk + 3) // 4
    assert nslices == min(24, nch)
    rounds, tail = divmod(ntiles, npairs)
    assert per_pair.min() >= rounds * nk
    if tail == 0:
        assert (per_pair == rounds * nk).all()
    else:
        per_tile = npairs // tail
        assert split_tiles == (tail if per_tile > 1 else 0)
        # a pair's share of the tail: at most ceil(nslices / per_tile) slices of at most ceil(nch / nslices) chunks
        worst = -(-nslices // max(per_tile, 1)) * -(-nch // nslices) * 4
        assert per_pair.max() <= rounds * nk + min(nk, worst)


def test_gram_plan_tail_is_spread_over_the_idle_pairs(L):
    # 820 tiles on 74 pairs: 11 full rounds + 6 tail tiles, each cut over 12 pairs (2 of 24 slices each)
    nk = 31250
    _, per_pair, nslices, split_tiles = plan(L, 820, 74, nk)
    assert nslices == 24 and split_tiles == 6
    busy = per_pair[per_pair > 11 * nk]
    assert len(busy) == 72
    assert busy.max() - 11 * nk <= nk // 12 + 8
    # the whole launch costs 11.09 tile-times instead of 12
    assert per_pair.max() / nk < 11.1


def schedule(L, p, pw):
    buf = np.zeros(4096, dtype=np.int64)
    npan = L.b200admm_k_panel_schedule(p, pw, buf.ctypes.data, len(buf))
    assert npan >= 1
    return buf[: npan + 1]


@pytest.mark.parametrize("p,pw", [(10000, 768), (10000, 256), (1100, 256), (1024, 768), (80000, 6000), (4352, 1000), (300, 768), (257, 100)])
def test_panel_schedule(L, p, pw):
    b = schedule(L, p, pw)
    assert b[0] == 0 and b[-1] == p
    assert (np.diff(b) > 0).all()
    assert (b[1:-1] % 256 == 0).all()                      # every interior boundary is a whole 256-column row block
    w = max(256, pw // 256 * 256)
    assert np.diff(b).max() <= w
    if p > w + 768:
        assert np.diff(b)[-1] <= 256 and np.diff(b)[-2] <= 256      # the tail of the copy is followed by little Gram work


def test_panel_schedule_rejects_bad_arguments(L):
    buf = np.zeros(4, dtype=np.int64)
    assert L.b200admm_k_panel_schedule(0, 256, buf.ctypes.data, 4) == -1
    assert L.b200admm_k_panel_schedule(100000, 256, buf.ctypes.data, 4) == -1      # not enough room


def eigen_linspaced(n, low, high):
    """Eigen 3.3 DenseBase::setLinSpaced(n, low, high) for floating point (NullaryFunctors.h, linspaced_op):
    n == 1 -> high; |high| < |low| -> first = low, rest = high - (n-1-i) step; else low + i step, last = high."""
    if n == 1:
        return np.array([high])
    step = (high - low) / (n - 1)
    i = np.arange(n, dtype=np.float64)
    if abs(high) < abs(low):
        v = high - (n - 1 - i) * step
        v[0] = low
    else:
        v = low + i * step
        v[-1] = high
    return v


@pytest.mark.parametrize("lmax,ratio,nl", [(0.37, 1e-4, 100), (512.0, 1e-2, 100), (3.0, 1e-4, 2), (0.9, 0.01, 1), (40.0, 1e-4, 1), (1.7, 0.5, 7)])
def test_lambda_grid_has_eigen_linspaced_semantics(L, lmax, ratio, nl):
    """src/Lasso.cpp:86-88.  nlambda = 1 must give lmin_ratio * lmax (setLinSpaced(1, low, high) is `high`),
    in the library and in the oracle alike."""
    out = np.zeros(nl)
    assert L.b200admm_k_lambda_grid(lmax, ratio, nl, out.ctypes.data) == 0
    import math
    want = np.array([math.exp(v) for v in eigen_linspaced(nl, math.log(lmax), math.log(ratio * lmax))])   # libm, as std::exp
    assert np.array_equal(out, want), np.abs(out / want - 1).max()
    if nl == 1:
        assert out[0] == math.exp(math.log(ratio * lmax)) and out[0] < lmax


def test_oracle_single_lambda_default_is_the_low_end():
    from oracle import pyoracle as O
    rng = np.random.default_rng(4)
    x = np.asfortranarray(rng.normal(size=(60, 8)))
    y = x[:, 0] - 0.5 * x[:, 1] + 0.1 * rng.normal(size=60)
    one = O.lasso_path(x, y, nlambda=1)
    two = O.lasso_path(x, y, nlambda=2)
    assert np.isclose(one["lambda_"][0], two["lambda_"][1], rtol=1e-14)
    assert np.isclose(one["lambda_"][0], 1e-4 * two["lambda_"][0], rtol=1e-12)


@pytest.mark.parametrize("p", [3, 8, 20, 37, 256, 1001, 1100, 2310, 4100, 10000, 10001, 20000])
@pytest.mark.parametrize("sms", [148, 132, 16])
def test_one_triangle_row_assignment_covers_every_row_once_and_balances_the_bytes(L, p, sms):
    """tall_path_tri_kernel reads row i of the symmetric K^-1 up to the diagonal (i + 1 entries); CTA c owns the folded
    rows [t0, t1) and [p - t1, p - t0): every row exactly once, and -- a folded pair has p + 1 entries -- the same number
    of entries per CTA up to one pair (the last CTA may own fewer)."""
    rows = np.zeros(4 * 1024, dtype=np.int32)
    smem = C.c_longlong(0)
    G = L.b200admm_k_tri_plan(p, sms, rows.ctypes.data, 1024, C.byref(smem))
    if G == 0:                          # few SMs: the per-warp row-sum slots of 2 x 625 rows do not fit next to 2 p floats
        assert sms == 16 and p == 20000
        return
    assert 0 < G <= sms and smem.value <= 227 * 1024
    seen = np.zeros(p, dtype=np.int32)
    work = []
    for c in range(G):
        t0, t1, b0, b1 = (int(v) for v in rows[4 * c:4 * c + 4])
        assert 0 <= t0 <= t1 <= b0 <= b1 <= p
        seen[t0:t1] += 1
        seen[b0:b1] += 1
        work.append(sum(i + 1 for i in range(t0, t1)) + sum(i + 1 for i in range(b0, b1)))
    assert (seen == 1).all()
    assert sum(work) == p * (p + 1) // 2
    full = work[:-1] if G > 1 else work
    assert max(full) - min(full) <= p + 1                       # all but the last CTA: within one folded pair of each other
    assert work[-1] <= max(full)


def test_one_triangle_kernel_is_declined_when_two_p_vectors_do_not_fit_shared_memory(L):
    assert L.b200admm_k_tri_plan(60000, 148, None, 0, None) == 0
    assert L.b200admm_k_tri_plan(10000, 148, None, 0, None) == 148
