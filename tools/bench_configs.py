"""Secondary BASELINE.json configurations and kernel micro-benchmarks (run on the GPU box).

    python tools/bench_configs.py zu | enet | lad | bp | wide [options]

Each prints one JSON line.  Sizes default to BASELINE.json's (C3 / C4); the CPU
oracle is NOT run here (bench.py holds the headline CPU comparison).
"""
import argparse
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import admm_b200
from admm_b200 import _capi as K

HBM_PEAK = 6552.6
try:
    HBM_PEAK = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass


def synth(n, p, seed=123, nsig=100, noise=1.0, mean=0.0):
    X = torch.empty((p, n), dtype=torch.float32, device="cuda")
    y = torch.empty(n, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_synth_f32(X.data_ptr(), y.data_ptr(), n, p, 0, seed, mean, 2.0, min(nsig, p), noise))
    return X, y


def zu(args):
    """fused z + u + residual + norms kernel on vectors far larger than L2 (the >= 90 % HBM target)."""
    L = args.len
    g = torch.Generator(device="cuda").manual_seed(1)
    x, ay, oz, az = (torch.randn(L, device="cuda", generator=g) for _ in range(4))
    z = torch.empty(L, device="cuda"); y = torch.empty(L, device="cuda")
    sums = np.zeros(6); ms = C.c_float(0)
    torch.cuda.synchronize()
    K.check(K.lib().b200admm_k_fused_zu_f32(x.data_ptr(), ay.data_ptr(), oz.data_ptr(), az.data_ptr(), z.data_ptr(), y.data_ptr(),
                                            L, 0.7, 1.3, 0, 1.0, sums.ctypes.data, C.byref(ms), args.repeats))
    gbs = 6 * 4 * L / (ms.value * 1e-3) / 1e9
    return {"bench": "fused_zu_f32", "len": L, "ms": ms.value, "algorithmic_bytes": 24 * L, "achieved_gbs": gbs,
            "peak_gbs": HBM_PEAK, "frac": gbs / HBM_PEAK, "note": "reads x, adj_y, old_z, adj_z; writes z, y; 6 L B bytes; vectors >> L2"}


def enet(args):
    n, p = args.n or 500_000, args.p or 5_000
    X, y = synth(n, p)
    f0 = admm_b200.admm_enet(X.t(), y).penalty(nlambda=2, alpha=0.5).fit()     # lambda_max from a first call
    lam = [0.1 * float(f0.lambda_[0])]
    admm_b200.admm_enet(X.t(), y).penalty(lam, alpha=0.5).fit()
    t0 = time.perf_counter()
    f = admm_b200.admm_enet(X.t(), y).penalty(lam, alpha=0.5).fit()
    wall = time.perf_counter() - t0
    T = f.info["timing"]; it = int(f.niter.sum())
    bpi = 4.0 * p * (p + 1) + 64.0 * p
    return {"bench": "enet_tall", "n": n, "p": p, "alpha": 0.5, "lambda": lam[0], "niter": it, "wall_s": wall, "phase_s": T,
            "iters_per_s": it / T["iterate"], "iter_gbs": bpi * it / T["iterate"] / 1e9, "hbm_frac": bpi * it / T["iterate"] / 1e9 / HBM_PEAK,
            "nnz": int(f.beta.nnz), "note": "K^-1 is 100 MB: L2-resident, so the fraction exceeds the HBM roofline"}


def lad(args):
    n, p = args.n or 500_000, args.p or 5_000
    X, y = synth(n, p, noise=1.0, mean=0.0)
    Xd = X.double(); yd = y.double()
    del X
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    f = admm_b200.admm_lad(Xd.t(), yd).opts(maxit=args.maxit).fit()
    wall = time.perf_counter() - t0
    T = f.info["timing"]
    bpi = 2.0 * n * p * 8 + p * p * 8 + 16.0 * n * 8
    return {"bench": "lad", "n": n, "p": p, "niter": f.niter, "wall_s": wall, "phase_s": T, "ms_per_iter": T["iterate"] / f.niter * 1e3,
            "iter_gbs": bpi * f.niter / T["iterate"] / 1e9, "hbm_frac": bpi * f.niter / T["iterate"] / 1e9 / HBM_PEAK,
            "gram_tflops_f64": n * p * (p + 1) / T["gram"] / 1e12, "rho_final": f.info["rho"]}


def bp(args):
    n, p = args.n or 5_000, args.p or 500_000
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn((p, n), device="cuda", dtype=torch.float64, generator=g)          # column-major n x p
    bt = torch.zeros(p, device="cuda", dtype=torch.float64)
    idx = torch.randperm(p, device="cuda", generator=g)[:500]
    bt[idx] = torch.rand(500, device="cuda", dtype=torch.float64, generator=g)
    b = A.t() @ bt
    t0 = time.perf_counter()
    f = admm_b200.admm_bp(A.t(), b).opts(maxit=args.maxit).fit()
    wall = time.perf_counter() - t0
    T = f.info["timing"]
    beta = torch.from_numpy(np.asarray(f.beta.todense())[:, 0]).cuda()
    err = float((beta - bt).abs().max())
    bpi = 2.0 * n * p * 8 + 16.0 * p * 8
    return {"bench": "bp", "n": n, "p": p, "nsig": 500, "niter": f.niter, "wall_s": wall, "phase_s": T, "ms_per_iter": T["iterate"] / f.niter * 1e3,
            "iter_gbs": bpi * f.niter / T["iterate"] / 1e9, "hbm_frac": bpi * f.niter / T["iterate"] / 1e9 / HBM_PEAK,
            "recovery_max_err": err, "rho_final": f.info["rho"]}


def wide(args):
    n, p = args.n or 10_000, args.p or 1_000_000
    X, y = synth(n, p)
    t0 = time.perf_counter()
    f = admm_b200.admm_lasso(X.t(), y).penalty(nlambda=args.nlambda).opts(maxit=args.maxit).fit()
    wall = time.perf_counter() - t0
    T = f.info["timing"]
    return {"bench": "lasso_wide", "n": n, "p": p, "nlambda": args.nlambda, "niter_total": int(f.niter.sum()), "niter": f.niter.tolist(),
            "wall_s": wall, "phase_s": T, "us_per_iter": T["iterate"] / max(1, int(f.niter.sum())) * 1e6,
            "nnz_last": int(f.beta[:, -1].nnz), "gamma": f.info["eig"], "rho_final": f.info["rho"]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["zu", "enet", "lad", "bp", "wide"])
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--p", type=int, default=0)
    ap.add_argument("--len", type=int, default=1 << 28)
    ap.add_argument("--repeats", type=int, default=10)
    ap.add_argument("--maxit", type=int, default=10000)
    ap.add_argument("--nlambda", type=int, default=100)
    a = ap.parse_args()
    out = {"zu": zu, "enet": enet, "lad": lad, "bp": bp, "wide": wide}[a.which](a)
    out["device"] = admm_b200.device_info()["name"]
    print(json.dumps(out), flush=True)
