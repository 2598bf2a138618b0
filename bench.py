#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: ADMM iterations/sec and full lambda-path wall time for
Lasso n = 1e6 x p = 1e4 (100-lambda path, fp32, ADMMLassoTall) on 1 / 2 / 4 / 8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one complete admm_lasso(x, y)$penalty(nlambda = 100)$fit(): DataStd, X'y, the Gram
matrix, the coarse-Lanczos rho, the factorisation / inverse and all ADMM iterations of the
100-lambda warm-started path, through the library's C ABI (b200admm_lasso).

  value     whole-job ADMM iterations per second (sum of niter over the path * K / timed seconds),
            X already resident in HBM (float32, column-major) when the timed region starts;
  e2e       the same with X, y in pinned HOST memory: the host->device copy of the 40 GB design
            and the device->host read of the solutions are inside every timed step;
  roofline  the dominant kernel of the step, the tensor-core Gram kernel (gram_tc.cu): algorithmic flops
            n p (p + 1) / its device time against the measured sustained dense bf16 rate (`executed` = the 3.2x
            it really executes: three fp16 products per element); roofline_iteration: the persistent iteration
            kernel (fadmm_tall.cu), algorithmic bytes per iteration 4 p (p + 1) + 64 p / device time per
            iteration against measured HBM peak; `traffic` of both from profiles/traffic.json (ncu captures);
  cpu_baseline  the CPU oracle (restated reference, OpenBLAS on all host cores) on a bounded sample.

N > 1 (torchrun, one rank per GPU): the same n x p problem row-sharded over the ranks (strong
scaling): global standardisation / X'y / Gram by NCCL all-reduce, factorisation replicated, iterations
sharded over the rows of K^-1 with the exchange fused into the kernel over NVLink peer memory.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--p", type=int, default=10_000)
    ap.add_argument("--nlambda", type=int, default=100)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-rows", type=int, default=20_000, help="rows of the CPU Gram sample")
    ap.add_argument("--cpu-lambdas", type=int, default=12, help="lambdas of the CPU iteration sample")
    ap.add_argument("--seed", type=int, default=123)
    return ap.parse_args()


def peaks():
    """(HBM GB/s, dense bf16 TFLOP/s sustained, source).  The tensor-bound kernel is timed inside a long step,
    so the sustained bf16 figure is its denominator (B200_PROFILING.md)."""
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        d = json.load(open(f))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1350.0))), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1350.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, n, p):
    """dram__bytes_read + write of one launch from the committed `ncu --set full` capture (profiles/traffic.json),
    or None when no capture of this kernel at this size is on file."""
    try:
        for e in json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["captures"]:
            if e["kernel"] == kernel and e["n"] == n and e["p"] == p:
                return e
    except Exception:
        pass
    return None


class ClockSampler:
    """SM clock / throttle reasons during the timed region (the recipe's clocks line).  Sampled through NVML
    in-process (nvidia_ml_py); spawning nvidia-smi five times a second costs ~0.2-0.7 s of driver
    initialisation per call on an 8-GPU node and was seen to stall the timed CUDA calls themselves.
    nvidia-smi is the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()
        self.th = None

    def _nvml_handle(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                ent = vis.split(",")[self.index].strip()
                if ent.isdigit():
                    phys = int(ent)
                else:
                    return pynvml, pynvml.nvmlDeviceGetHandleByUUID(ent.encode() if hasattr(ent, "encode") else ent)
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:
            return None, None

    def _run_nvml(self, nv, h):
        bits = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8),
                ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append([str(sm), str(mx), str(pw)] + ["Active" if mask & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def _run(self):
        nv, h = self._nvml_handle()
        if nv is not None and h is not None:
            self.source = "nvml"
            return self._run_nvml(nv, h)
        self.source = "nvidia-smi"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag.set()
        if self.th:
            self.th.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "source": getattr(self, "source", None)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def algorithmic_bytes_per_iter(p):
    # factor read forward + backward (2 * p(p+1)/2 floats) + 16 vector passes
    return 4.0 * p * (p + 1) + 64.0 * p


# -------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference) on a bounded sample of the same workload
# -------------------------------------------------------------------------------------------------
def wishart_problem(n, p, seed):
    """(lower(X'X), X'y, lambda grid) of the standardised full-size problem WITHOUT forming X: for i.i.d. Gaussian
    columns the standardised Gram matrix is n on the diagonal and sqrt(n) N(0,1) off it (central limit of n = 1e6
    products), and X'y = (G 2 beta* + sqrt(n) N(0,1)) / sd(y) for y = X beta* + noise with X ~ N(0, 2^2).  The CPU
    iteration count on it is that of the real design (1063 +- a few at n = 1e6, p = 1e4), which a Gram matrix
    computed from a 20 000-row sample is not: its spectrum is far wider (p / n = 0.5 instead of 0.01)."""
    rng = np.random.default_rng(seed + 1)
    G = np.empty((p, p), dtype=np.float32, order="F")
    rt = np.float32(np.sqrt(n))
    for j0 in range(0, p, 1024):
        G[:, j0:j0 + 1024] = rng.standard_normal((p, min(1024, p - j0)), dtype=np.float32) * rt
    G[np.diag_indices(p)] = np.float32(n)                    # only the lower triangle is read
    m = min(100, p)
    b2 = np.zeros(p, dtype=np.float32)
    b2[:m] = 2.0 * rng.uniform(size=m)
    sdy = float(np.sqrt(1.0 + float((b2.astype(np.float64) ** 2).sum())))
    Gb = np.zeros(p, dtype=np.float64)
    L = np.tril(G[:, :m].astype(np.float64), -1)             # G b over the first m columns, symmetric completion
    Gb += L @ b2[:m].astype(np.float64)
    Gb[:m] += np.tril(G[:m, :m].astype(np.float64), -1).T @ b2[:m].astype(np.float64) + float(n) * b2[:m]
    xy = ((Gb + np.sqrt(n) * rng.standard_normal(p)) / sdy).astype(np.float32)
    return G, xy


def cpu_sample(args, niter_total=None):
    """Times the CPU path piecewise and extrapolates to the full configuration:
    DataStd + Gram on `cpu_rows` rows at full p (linear in n); then Lanczos + Cholesky and the ADMM
    iterations at full p.  GPU arm (niter_total given): iterations timed on the first `cpu_lambdas`
    lambdas of the sample's own Gram matrix and the per-iteration cost (which depends only on p) scaled
    to the GPU run's iteration count.  Reference arm (niter_total None): the whole 100-lambda path is run
    on a Gram matrix with the full-size problem's statistics (wishart_problem), so its iteration count
    and iteration seconds are measured, not scaled."""
    from oracle import pyoracle as O
    cores = host_threads()
    bt = O.use_openblas(cores)
    O.omp_threads(cores)
    n, p, nl = args.n, args.p, args.nlambda
    ns = min(args.cpu_rows, n)
    rng = np.random.default_rng(args.seed)
    x = np.empty((ns, p), dtype=np.float32, order="F")
    for j0 in range(0, p, 512):
        x[:, j0:j0 + 512] = rng.standard_normal((ns, min(512, p - j0)), dtype=np.float32) * 2.0
    beta = np.zeros(p, dtype=np.float32)
    beta[:100] = rng.uniform(size=min(100, p))[: min(100, p)] if p >= 100 else 0
    y = (x[:, :100] @ beta[:100] + rng.standard_normal(ns, dtype=np.float32)).astype(np.float32)
    t0 = time.perf_counter()
    st = O.standardize_f32(x, y)
    t_std = time.perf_counter() - t0
    t0 = time.perf_counter()
    G = O.gram_tn_f32(x)
    xy = (x.T @ y).astype(np.float32)
    t_gram = time.perf_counter() - t0
    full_path = niter_total is None
    if full_path:
        del G, xy
        G, xy = wishart_problem(n, p, args.seed)
        ncpu_l = nl
    else:
        # scale the sample Gram to the full problem's magnitude so rho / conditioning are comparable
        G *= np.float32(n / ns)
        xy *= np.float32(n / ns)
        ncpu_l = max(2, min(args.cpu_lambdas, nl))
    lam0 = float(np.abs(xy).max())
    grid = np.exp(np.linspace(np.log(lam0), np.log(lam0 * 1e-4), nl))[:ncpu_l]
    t0 = time.perf_counter()
    r = O.tall_path_from_gram(G, xy, grid)
    t_path = time.perf_counter() - t0
    t_setup = float(r["setup_s"])
    it_sample = int(r["niter"].sum())
    t_iter = max(t_path - t_setup, 1e-9)
    per_iter = t_iter / max(it_sample, 1)
    full_gram = t_gram * (n / ns)
    full_std = t_std * (n / ns)
    nit = it_sample if full_path else niter_total
    full_wall = full_std + full_gram + t_setup + per_iter * nit
    return {
        "value": nit / full_wall, "unit": "ADMM iters/s (whole lambda path incl. setup)", "cores": cores,
        "kind": "port",
        "sample": ("oracle (restated reference, OpenBLAS %d threads): DataStd+Gram timed on %d of %d rows at p=%d and scaled "
                   "linearly in n; Lanczos+Cholesky at full p; " % (bt, ns, n, p)) +
                  (("all %d lambdas run on a Gram matrix with the full-size design's statistics: %d iterations measured, "
                    "%.2f ms/iter" % (nl, it_sample, per_iter * 1e3)) if full_path else
                   ("iterations timed on the first %d of %d lambdas (%d iterations, %.2f ms/iter) and scaled to the GPU run's "
                    "%d iterations" % (ncpu_l, nl, it_sample, per_iter * 1e3, nit))),
        "path_wall_s_extrapolated": full_wall, "iters_per_s_steady": 1.0 / per_iter,
        "measured_s": {"standardize": t_std, "gram": t_gram, "lanczos_cholesky": t_setup, "iterations": t_iter},
    }


def run_reference(args, rank):
    if rank != 0:
        return
    t_all = time.perf_counter()
    cb = cpu_sample(args)      # one bounded sample (tens of seconds of CPU work) stands for every step
    line = {
        "impl": "reference", "metric": "admm_iters_per_sec_full_lambda_path", "value": cb["value"],
        "unit": "ADMM iters/s (whole lambda path incl. setup)", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cb["path_wall_s_extrapolated"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "lasso_tall_n%d_p%d_%dlambda" % (args.n, args.p, args.nlambda), "n": args.n, "p": args.p,
                   "nlambda": args.nlambda, "note": "bounded CPU sample, extrapolated piecewise (see cpu_baseline.sample)"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import admm_b200
    from admm_b200 import _capi as K

    torch.cuda.set_device(local_rank)
    L = K.lib()
    info = admm_b200.device_info()
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        idbuf = torch.zeros(K.COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            raw = C.create_string_buffer(K.COMM_ID_BYTES)
            K.check(L.b200admm_comm_id(raw))
            idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        idbuf = idbuf.cuda()
        dist.broadcast(idbuf, 0)
        raw = bytes(idbuf.cpu().numpy().tobytes())
        K.check(L.b200admm_comm_init(raw, rank, world))

    n, p, nl = args.n, args.p, args.nlambda
    chunk = n // world
    row0 = rank * chunk
    n_local = chunk if rank < world - 1 else n - row0

    # ---- synthetic design in HBM (README recipe: X ~ N(0, 2^2), 100 U(0,1) signals, unit noise) -----
    Xd = torch.empty((p, n_local), dtype=torch.float32, device="cuda")
    yd = torch.empty(n_local, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    K.check(L.b200admm_synth_f32(Xd.data_ptr(), yd.data_ptr(), n_local, p, row0, args.seed, 0.0, 2.0, min(100, p), 1.0))

    lib_stream = torch.cuda.ExternalStream(L.b200admm_stream())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fit_device():
        return admm_b200.admm_lasso(Xd.t(), yd).penalty(nlambda=nl).fit()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize, device time from CUDA events recorded on the
        library's own stream, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = L.b200admm_launch_count()
        t0 = time.perf_counter()
        e0.record(lib_stream)
        fits = [fn() for _ in range(steps)]
        e1.record(lib_stream)
        e1.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        dev = e0.elapsed_time(e1) * 1e-3
        if world > 1:
            t = torch.tensor([dev, wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dev, wall = float(t[0]), float(t[1])
        return fits, dev, wall, L.b200admm_launch_count() - launches0

    for _ in range(args.warmup):
        fit_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    fits, dev_s, wall_s, launches = timed(fit_device, args.steps)
    clocks = sampler.stop()

    niter_path = int(fits[-1].niter.sum())
    total_iters = sum(int(f.niter.sum()) for f in fits)
    T = {k: float(np.mean([f.info["timing"][k] for f in fits])) for k in fits[-1].info["timing"]}
    hbm_peak, tensor_peak, peak_src = peaks()
    bpi = algorithmic_bytes_per_iter(p)
    achieved = bpi * niter_path / T["iterate"] / 1e9
    gram_flops = float(n) * p * (p + 1)
    gram_kernel_s = float(L.b200admm_last_gram_seconds())         # last timed step, this rank's row block
    gram_alg_tflops = float(n_local) * p * (p + 1) / max(gram_kernel_s, 1e-9) / 1e12
    # executed tensor flops: three fp16 products per element, 256 x 256 tiles on and below the diagonal
    nb = (p + 255) // 256
    gram_exec_tflops = 3.0 * 2.0 * n_local * 65536.0 * (nb * (nb + 1) // 2) / max(gram_kernel_s, 1e-9) / 1e12
    tr_gram = ncu_traffic("gram_pair_h_kernel", n_local, p)
    tr_iter = ncu_traffic("tall_path_kernel", n, p)

    line = {
        "metric": "admm_iters_per_sec_full_lambda_path", "value": total_iters / dev_s,
        "unit": "ADMM iters/s (whole lambda path incl. setup)", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "lasso_tall_n%d_p%d_%dlambda" % (n, p, nl), "n": n, "p": p, "nlambda": nl,
                   "standardize": True, "intercept": True, "eps_abs": 1e-5, "eps_rel": 1e-5, "maxit": 10000, "rho": "auto",
                   "sharding": "rows over %d rank(s), Gram all-reduce, iterations sharded over peer memory" % world,
                   "l2": "inputs (%.1f GB) exceed L2; no flush needed" % (4.0 * n_local * p / 1e9)},
        "path_wall_s": dev_s / args.steps, "host_wall_s_per_step": wall_s / args.steps,
        "niter_path": niter_path, "iters_per_sec_steady": niter_path / T["iterate"],
        "phase_s": T,
        # the dominant kernel of the step (60 % of the device time): the Gram matrix on the tensor cores.
        # achieved = ALGORITHMIC flops n p (p + 1) (a symmetric rank-n update) per launch / kernel time; the kernel
        # executes 3.2x that (three fp16 products per element for fp32 accuracy, whole 256 x 256 tiles on the
        # diagonal), which `executed` states next to it.
        "roofline": {"kernel": "gram_pair_h_kernel (tcgen05 kind::f16 CTA-pair Gram, 3-product fp16 split)", "bound": "tensor",
                     "achieved": gram_alg_tflops, "peak": tensor_peak, "unit": "TFLOP/s", "frac": gram_alg_tflops / tensor_peak,
                     "traffic": (tr_gram["dram_bytes"] if tr_gram else None), "traffic_source": (tr_gram["source"] if tr_gram else None),
                     "peak_source": peak_src + ", dense bf16 sustained", "kernel_s": gram_kernel_s,
                     "algorithmic_flop": float(n_local) * p * (p + 1), "executed": gram_exec_tflops, "executed_frac": gram_exec_tflops / tensor_peak},
        # the per-iteration hot path named by BASELINE.json: one persistent launch for the whole lambda path
        "roofline_iteration": {"kernel": "tall_path_kernel (persistent lambda-path iteration kernel)", "bound": "hbm",
                               "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                               # DRAM bytes per iteration from the ncu capture (below the algorithmic figure: the
                               # alternating sweep direction leaves the tail of K^-1 in L2), scaled to this launch
                               "traffic": (tr_iter["dram_bytes_per_iteration"] * niter_path if tr_iter else None),
                               "traffic_source": (tr_iter["source"] if tr_iter else None),
                               "traffic_unit": "bytes per launch (one launch = the whole lambda path)", "peak_source": peak_src,
                               "bytes_per_iteration": bpi, "us_per_iteration": T["iterate"] / max(niter_path, 1) * 1e6},
        "setup_flops": {"gram_syrk_flop": gram_flops, "gram_phase_tflops": gram_flops / world / max(T["gram"], 1e-9) / 1e12},
        "clocks": clocks, "gpu_launches": int(launches), "device": info["name"],
    }

    # ---- end to end: host buffers in, host results out, every step -------------------------------
    if not args.no_e2e:
        try:
            Xh = torch.empty((p, n_local), dtype=torch.float32, pin_memory=True)
            yh = torch.empty(n_local, dtype=torch.float32, pin_memory=True)
            Xh.copy_(Xd); yh.copy_(yd)
            torch.cuda.synchronize()
            xh_np, yh_np = Xh.numpy().T, yh.numpy()

            def fit_host():
                return admm_b200.admm_lasso(xh_np, yh_np).penalty(nlambda=nl).fit()
            fit_host()
            fh, e_dev, e_wall, _ = timed(fit_host, args.e2e_steps)
            it = sum(int(f.niter.sum()) for f in fh)
            line["e2e"] = {"value": it / e_wall, "unit": line["unit"],
                           "h2d_bytes_per_step": int(4 * n_local * p + 4 * n_local),
                           "d2h_bytes_per_step": int(4 * nl * p + 4 * nl + 3 * 4 * p + 8),
                           "path_wall_s": e_wall / args.e2e_steps, "steps": args.e2e_steps,
                           "device_s_per_step": e_dev / args.e2e_steps,
                           "phase_s": {k: float(np.mean([f.info["timing"][k] for f in fh])) for k in fh[-1].info["timing"]},
                           "phase_s_per_step": [{k: round(float(v), 4) for k, v in f.info["timing"].items()} for f in fh],
                           "input": "float32 column-major, pinned host memory"}
            del Xh, yh
        except Exception as ex:  # pinned allocation can fail on small hosts: say so, do not fake a number
            line["e2e"] = {"value": None, "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                           "error": repr(ex)[:300]}

    if rank == 0 and world == 1 and not args.no_cpu:
        del Xd
        torch.cuda.empty_cache()
        try:
            line["cpu_baseline"] = cpu_sample(args, niter_total=niter_path)
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "error": repr(ex)[:300]}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        L.b200admm_comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
