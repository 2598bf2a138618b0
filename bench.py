#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: ADMM iterations/sec and full lambda-path wall time, measured through the
library's C ABI on 1 / 2 / 4 / 8 B200, with the reference's CPU path (the oracle restatement) timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--config tall|enet|wide|lad|bp|consensus]

--config tall (default, BASELINE.json configs[1], the configuration the metric is quoted on):
    Lasso n = 1e6 x p = 1e4, 100-lambda path, fp32 (ADMMLassoTall).  A "step" is one complete
    admm_lasso(x, y)$penalty(nlambda = 100)$fit(): DataStd, X'y, the Gram matrix, the coarse-Lanczos rho, the
    factorisation / inverse and all ADMM iterations of the warm-started path (b200admm_lasso).
      value     whole-job ADMM iterations per second (sum of niter over the path * K / timed seconds), X already
                resident in HBM (float32, column-major) when the timed region starts;
      e2e       the same with X, y in pinned HOST memory: the host->device copy of the 40 GB design and the
                device->host read of the solutions are inside every timed step;
      roofline  the dominant kernel of the step, the tensor-core Gram kernel; roofline_iteration: the persistent
                iteration kernel (HBM); `traffic` of both from profiles/traffic.json (ncu captures);
      parity    N = 1: the CPU oracle runs the whole 100-lambda path on the GPU's OWN Gram matrix / X'y / lambda grid
                (captured through b200admm_set_capture) and the coefficients, supports and iteration counts are
                compared; 10^4 random Gram entries are checked against float64 dot products of the generated columns;
                and, when the reference arm ran on this box before (same generator -> bit-identical X), the GPU fit is
                compared with the CPU fit of the full design.  N > 1 (`parity_vs_n1`): rank 0 also fits the undivided
                problem on one GPU and every rank's sharded result is compared with it (and with the other ranks',
                bitwise).  A miss makes the process exit with status 3 after printing the line;
      cpu_baseline  the oracle with OpenBLAS on all host cores: the iteration phase is the parity run above (measured,
                full size); DataStd + Gram on a 20 000-row sample of the same design, scaled linearly in n.
    --impl reference: the oracle fits the FULL design (n = 1e6 rows streamed through the bit-identical CPU generator,
    DataStd + chunked SYRK Gram + Lanczos + Cholesky + the 100-lambda path), everything measured; generation time is
    excluded (X is the caller's input).  One fit is timed whatever --steps says (`steps_measured`).

--config enet | wide | lad | bp | consensus: BASELINE.json configs[2..4] (C3, C4, C5) with the
same keys; see the functions below for each workload's parity check and roofline.

N > 1 (torchrun, one rank per GPU): tall / enet = the same n x p problem row-sharded over the ranks (strong scaling:
global standardisation / X'y / Gram by NCCL all-reduce, iterations sharded over the rows of K^-1 with the exchange
fused into the kernel over NVLink peer memory); consensus = the reference's $parallel(N), one row block per GPU and one
all-reduce per iteration; wide / lad / bp do not shard (replicas are not run: rank 0 alone measures).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "ADMM iters/s (whole lambda path incl. setup)"
UNIT_FIT = "ADMM iters/s (whole fit incl. setup)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="tall", choices=["tall", "enet", "wide", "lad", "bp", "consensus"])
    ap.add_argument("--n", "--rows", dest="n", type=int, default=0)       # (--rows / --cols: torchrun's own parser trips over "--n")
    ap.add_argument("--p", "--cols", dest="p", type=int, default=0)
    ap.add_argument("--nlambda", type=int, default=100)
    ap.add_argument("--maxit", type=int, default=10000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-rows", type=int, default=20_000, help="rows of the CPU DataStd + Gram sample (GPU arm)")
    ap.add_argument("--ref-sample", action="store_true", help="reference arm: bounded sample + extrapolation instead of the full fit")
    ap.add_argument("--seed", type=int, default=123)
    a = ap.parse_args()
    dn, dp = {"tall": (1_000_000, 10_000), "enet": (500_000, 5_000), "wide": (10_000, 1_000_000), "lad": (500_000, 5_000),
              "bp": (5_000, 500_000), "consensus": (1_000_000, 80_000 if a.gpus > 1 else 20_000)}[a.config]
    a.n = a.n or dn
    a.p = a.p or dp
    return a


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
        sm, mx, reasons, pw = [], 0.0, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1])); pw.append(float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_median": float(np.median(pw)) if pw else None,
                "source": getattr(self, "source", None)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def tall_bytes_per_iter(p):
    # factor read forward + backward (2 * p(p+1)/2 floats) + 16 vector passes
    return 4.0 * p * (p + 1) + 64.0 * p


def tri_kernel_in_use(world):
    """One GPU (and B200ADMM_TALL_TRI not 0): the iteration kernel reads one triangle of the symmetric K^-1."""
    return world == 1 and os.environ.get("B200ADMM_TALL_TRI", "1") != "0"


def iteration_roofline(env, p, niter_path, iterate_s, world, tr_iter):
    """The per-iteration hot path named by BASELINE.json: one persistent launch for the whole lambda path.  Two byte
    counts per iteration: SURVEY.md section 8(d)'s 4 p (p + 1) + 64 p (the reference's two triangular solves = one full
    read of K^-1) and 2 p (p + 1) + 64 p (one triangle of the symmetric K^-1).  `frac` is quoted on the bytes the
    kernel in use has to read: the triangle on one GPU (tall_path_tri_kernel), the full rows on row-sharded runs."""
    full = tall_bytes_per_iter(p)
    tri = 2.0 * p * (p + 1) + 64.0 * p
    use_tri = tri_kernel_in_use(world)
    ach_full = full * niter_path / iterate_s / 1e9
    ach_tri = tri * niter_path / iterate_s / 1e9
    ach = ach_tri if use_tri else ach_full
    return {"kernel": ("tall_path_tri_kernel (persistent lambda-path iteration kernel, one triangle of K^-1 per iteration)" if use_tri
                       else "tall_path_kernel (persistent lambda-path iteration kernel, full rows of K^-1)"),
            "bound": "hbm", "achieved": ach, "peak": env.hbm_peak, "unit": "GB/s", "frac": ach / env.hbm_peak,
            "bytes_per_iteration": tri if use_tri else full,
            "achieved_on_reference_bytes": ach_full, "frac_on_reference_bytes": ach_full / env.hbm_peak, "reference_bytes_per_iteration": full,
            "achieved_one_triangle": ach_tri,
            "traffic": (tr_iter["dram_bytes_per_iteration"] * niter_path if tr_iter else None),
            "traffic_source": (tr_iter["source"] if tr_iter else None),
            "traffic_unit": "bytes per launch (one launch = the whole lambda path)", "peak_source": env.peak_src,
            "us_per_iteration": iterate_s / max(niter_path, 1) * 1e6}


def ref_fit_file(args):
    return os.path.join(tempfile.gettempdir(), "b200admm_ref_%s_n%d_p%d_l%d_s%d.npz" % (args.config, args.n, args.p, args.nlambda, args.seed))


# ---------------------------------------------------------------------------------------------------------------
# host memory placement for the end-to-end arm
# ---------------------------------------------------------------------------------------------------------------
def bind_near_gpu(local_rank):
    """Place this rank's pinned staging memory on the NUMA node its GPU hangs off: CPU affinity to that node's
    cores (first-touch allocation follows the thread) and a preferred-node memory policy.  Best effort; returns
    what was done so the bench line can say where the buffers live."""
    info = {"gpu_numa_node": None, "cpus": None, "mempolicy": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis and vis.split(",")[local_rank].strip().isdigit():
            idx = int(vis.split(",")[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        cl = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        cpus = set()
        for part in cl.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            info["cpus"] = "%d of node %d's %d cores" % (len(use), node, len(cpus))
        else:
            info["cpus"] = "node %d's cores are outside this process's cpuset (%d allowed cpus)" % (node, len(allowed))
        libc = C.CDLL("libc.so.6", use_errno=True)
        mask = C.c_ulong(1 << node)
        rc = libc.syscall(238, 1, C.byref(mask), C.c_ulong(64))            # set_mempolicy(MPOL_PREFERRED, {node})
        info["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy failed (errno %d)" % C.get_errno()
    except Exception as ex:
        info["error"] = repr(ex)[:200]
    return info


# ---------------------------------------------------------------------------------------------------------------
# shared GPU-arm plumbing
# ---------------------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import admm_b200
        from admm_b200 import _capi as K
        self.args, self.torch, self.dist, self.A, self.K = args, torch, dist, admm_b200, K
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.L = K.lib()
        self.info = admm_b200.device_info()
        if self.world > 1:
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", self.local_rank))
            from admm_b200 import dist as D
            D.init_comm()
        self.lib_stream = torch.cuda.ExternalStream(self.L.b200admm_stream())
        self.hbm_peak, self.tensor_peak, self.peak_src = peaks()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K steps bracketed by barrier + synchronize, device time from CUDA events recorded on the library's own
        stream, max over ranks.  Returns (results, device seconds, wall seconds, kernel launches)."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = self.L.b200admm_launch_count()
        t0 = time.perf_counter()
        e0.record(self.lib_stream)
        out = [fn() for _ in range(steps)]
        e1.record(self.lib_stream)
        e1.synchronize()
        self.barrier()
        wall = time.perf_counter() - t0
        dev = e0.elapsed_time(e1) * 1e-3
        if self.world > 1:
            t = torch.tensor([dev, wall], dtype=torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            dev, wall = float(t[0]), float(t[1])
        return out, dev, wall, self.L.b200admm_launch_count() - launches0

    def synth(self, n_rows, p, row0, nsig=100, noise=1.0, mean=0.0, sd=2.0):
        torch = self.torch
        X = torch.empty((p, n_rows), dtype=torch.float32, device="cuda")
        y = torch.empty(n_rows, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        self.K.check(self.L.b200admm_synth_f32(X.data_ptr(), y.data_ptr(), n_rows, p, row0, self.args.seed, mean, sd, min(nsig, p), noise))
        return X, y

    def finish(self):
        if self.world > 1:
            self.L.b200admm_comm_destroy()
            self.dist.destroy_process_group()

    def base_line(self, metric, unit, value, dev_s, scaling, dtype, config):
        a = self.args
        return {"metric": metric, "value": value, "unit": unit, "n_gpus": self.world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": dev_s / a.steps * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": dtype, "data": "synthetic", "config": config}


def recover_f32(z, meanX, scaleX, meanY, scaleY):
    """DataStd::recover, flag 3, in float (DataStd.h:183-207) for a dense standardised coefficient vector."""
    z = z.astype(np.float32)
    nz = z != 0
    c = np.zeros_like(z)
    c[nz] = (z[nz] / scaleX[nz]) * np.float32(scaleY)
    b0 = np.float32(meanY) - np.float32((c[nz] * meanX[nz]).sum(dtype=np.float32))
    return np.concatenate([[np.float64(b0)], c.astype(np.float64)])


def compare_paths(bg, bc, ng, nc, lam_std, tol_rel, band_rel, p=None, eps_abs=1e-5):
    """Coefficient / support / iteration-count comparison of two lambda paths (columns = lambdas; row 0 = intercept).

    Two bounds, because the reference's stopping rule is coarse: it accepts any iterate whose residual norms are below
    sqrt(p) eps_abs + eps_rel |x| (eps = 1e-5), and near the small-lambda end of a warm-started path the iterate creeps
    along at that level for many iterations -- a last-bit difference in K^-1 (explicit inverse here, LLT solve in the
    reference) or in a norm moves the stopping iteration of single lambdas by up to a dozen iterations while the total
    stays put (measured at n = 1e6, p = 1e4: 66 to 96 of 100 lambdas stop at the identical iteration depending on which
    factorisation kernels built K^-1; the others differ by <= 12).
      * lambdas where both runs stop at the SAME iteration: max|dbeta| <= tol_rel * max(1, |beta|_inf)   (iterate-level parity)
      * all lambdas: max|dbeta| <= (tol_rel + 2 sqrt(p) eps_abs) * max(1, |beta|_inf)                  (what the stopping rule leaves open)
      * supports identical outside a band of band_rel * max(1, |beta|_inf); iteration totals within 3 %; at least half of
        the lambdas with identical iteration counts."""
    scale = max(1.0, float(np.abs(bc).max()))
    p = p if p is not None else bg.shape[0] - 1
    d = np.abs(bg - bc)
    mism = (bg[1:] != 0) != (bc[1:] != 0)
    big = np.maximum(np.abs(bg[1:]), np.abs(bc[1:]))
    outside = int((mism & (big > band_rel * scale)).sum())
    dn = np.abs(ng.astype(int) - nc.astype(int))
    eq = dn == 0
    d_eq = float(d[:, eq].max()) if eq.any() else 0.0
    loose = tol_rel + 2.0 * np.sqrt(p) * eps_abs
    res = {"max_abs_dbeta": float(d.max()), "max_abs_dbeta_where_niter_equal": d_eq, "max_abs_dintercept": float(d[0].max()),
           "beta_inf": float(np.abs(bc).max()), "tol_where_niter_equal": tol_rel * scale, "tol_all_lambdas": loose * scale,
           "support_size_last_lambda": [int((bg[1:, -1] != 0).sum()), int((bc[1:, -1] != 0).sum())],
           "support_mismatch": int(mism.sum()), "support_mismatch_outside_band": outside, "band": band_rel * scale,
           "largest_mismatched_coef": float(big[mism].max()) if mism.any() else 0.0,
           "niter_gpu": int(ng.sum()), "niter_cpu": int(nc.sum()), "niter_max_abs_diff_per_lambda": int(dn.max()),
           "lambdas_with_equal_niter": int(eq.sum()), "nlambda": int(len(ng))}
    res["ok"] = bool(d_eq <= tol_rel * scale and d.max() <= loose * scale and outside == 0
                     and abs(int(ng.sum()) - int(nc.sum())) <= max(3, 0.03 * int(nc.sum())) and 2 * int(eq.sum()) >= len(ng))
    return res


def trace_rel_diff(tg, tc):
    """Largest relative difference of two iteration traces; entries that are exactly zero in both (the dual residual
    of a first iteration) count as equal."""
    tg, tc = np.asarray(tg, dtype=np.float64), np.asarray(tc, dtype=np.float64)
    d = np.abs(tg - tc) / np.maximum(np.abs(tc), 1e-300)
    d[(tg == 0) & (tc == 0)] = 0.0
    return d.max() if d.size else 0.0


def align_traces(ta, tb, rtol):
    """Align two FADMM iteration traces (rows = eps_primal, resid_primal, eps_dual, resid_dual, rho) modulo "stutter"
    rows: an iteration that took the restart branch while z stood still repeats the previous one (dual residual exactly
    0, primal residual unchanged) and the run continues as the other run shifted by one row.  Returns the number of
    leading rows equal in both (`prefix`), matched row pairs, skipped (stutter) rows of either trace, the largest
    relative difference over the matched pairs, whether all of tb was consumed, and the index of ta's last matched row."""
    ta, tb = np.asarray(ta, dtype=np.float64), np.asarray(tb, dtype=np.float64)

    def stutter(t, k):
        return k > 0 and t[k][3] == 0.0 and abs(t[k][1] - t[k - 1][1]) <= rtol * abs(t[k - 1][1])
    i = j = matched = 0
    prefix, in_prefix, max_rel, last_a = 0, True, 0.0, -1
    sk_a, sk_b = [], []
    while i < len(ta) and j < len(tb):
        d = float(trace_rel_diff(ta[i], tb[j]))
        if d <= rtol:
            matched += 1
            max_rel = max(max_rel, d)
            last_a = i
            if in_prefix and i == j:
                prefix += 1
            i += 1
            j += 1
            continue
        in_prefix = False
        if stutter(ta, i):
            sk_a.append(i)
            i += 1
        elif stutter(tb, j):
            sk_b.append(j)
            j += 1
        else:
            break
    while j < len(tb) and stutter(tb, j):                 # trailing stutter rows of tb leave its state where it was
        sk_b.append(j)
        j += 1
    return {"prefix": prefix, "matched": matched, "skipped_a": sk_a, "skipped_b": sk_b, "max_rel": max_rel, "ok": bool(j == len(tb)),
            "last_a": last_a}


def gram_spot_check(env, Xd, n_local, cap, k=100):
    """k x k Gram entries of the captured matrix against float64 dot products of the standardised generated columns
    (torch float64 on the GPU as the checker).  Single-rank only (the columns must be whole)."""
    torch = env.torch
    p = Xd.shape[0]
    g = torch.Generator(device="cpu").manual_seed(7)
    I = torch.randperm(p, generator=g)[:k].sort().values
    J = torch.randperm(p, generator=g)[:k].sort().values
    def std_cols(idx):
        Z = Xd[idx.cuda()].double()                       # k x n
        Z = Z - Z.mean(dim=1, keepdim=True)
        return Z / (Z.norm(dim=1, keepdim=True) / np.sqrt(n_local))
    Zi, Zj = std_cols(I), std_cols(J)
    ref = (Zi @ Zj.t()).cpu().numpy()
    got = cap.gram[np.ix_(I.numpy(), J.numpy())].astype(np.float64)
    err = np.abs(got - ref)
    return {"entries": int(k * k), "max_abs_err": float(err.max()), "max_err_over_n": float(err.max() / n_local),
            "rms_err_over_n": float(np.sqrt((err ** 2).mean()) / n_local), "symmetric": bool(np.array_equal(cap.gram, cap.gram.T))}


def oracle_on_captured_gram(O, cap, fit, n, enet=False, alpha=1.0):
    """The CPU oracle's lambda path on the GPU's own (G, X'y, lambda grid); its rho comes from its own Lanczos run
    on G.  Returns (beta_cpu[(p+1) x nl], result dict, seconds of the iteration phase, setup seconds)."""
    ilam = fit.lambda_ * float(n) / float(cap.scaleY)
    t0 = time.perf_counter()
    o = O.tall_path_from_gram(cap.gram, cap.xy, ilam, enet=enet, alpha=alpha)
    t_all = time.perf_counter() - t0
    bc = np.stack([recover_f32(o["z"][k], cap.meanX, cap.scaleX, cap.meanY, cap.scaleY) for k in range(len(ilam))], axis=1)
    return bc, o, max(t_all - float(o["setup_s"]), 1e-9), float(o["setup_s"])


def cpu_std_gram_sample(O, args, ns):
    """DataStd + Gram + X'y of the oracle on the first ns rows of the same synthetic design (timed; linear in n)."""
    X, y = O.synth_f32(ns, args.p, row0=0, seed=args.seed, nsig=min(100, args.p))
    t0 = time.perf_counter()
    O.standardize_f32(X, y)
    t_std = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.gram_tn_f32(X)
    (X.T @ y)
    t_gram = time.perf_counter() - t0
    return t_std, t_gram


# ---------------------------------------------------------------------------------------------------------------
# --config tall / enet : n > p lasso / elastic-net path
# ---------------------------------------------------------------------------------------------------------------
def run_tall(args, enet=False):
    env = Env(args)
    torch, dist, A, K, L = env.torch, env.dist, env.A, env.K, env.L
    rank, world = env.rank, env.world
    n, p = args.n, args.p
    alpha = 0.5
    single = enet                                            # C4: one lambda = 0.1 lambda_max
    nl = 1 if single else args.nlambda
    chunk = n // world
    row0 = rank * chunk
    n_local = chunk if rank < world - 1 else n - row0
    Xd, yd = env.synth(n_local, p, row0)

    def model(x, y):
        return A.admm_enet(x, y) if enet else A.admm_lasso(x, y)

    lam_user = None
    if single:
        f0 = model(Xd.t(), yd).penalty(nlambda=2, alpha=alpha).opts(maxit=1).fit()
        lam_user = [0.1 * float(f0.lambda_[0])]

    def pen(m):
        if enet:
            return m.penalty(lam_user, alpha=alpha)
        return m.penalty(nlambda=nl)

    def fit_device():
        return pen(model(Xd.t(), yd)).fit()

    for _ in range(args.warmup):
        fit_device()
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    fits, dev_s, wall_s, launches = env.timed(fit_device, args.steps)
    clocks = sampler.stop()

    niter_path = int(fits[-1].niter.sum())
    total_iters = sum(int(f.niter.sum()) for f in fits)
    T = {k: float(np.mean([f.info["timing"][k] for f in fits])) for k in fits[-1].info["timing"]}
    bpi = tall_bytes_per_iter(p)
    achieved = bpi * niter_path / T["iterate"] / 1e9
    gram_flops = float(n) * p * (p + 1)
    gram_kernel_s = float(L.b200admm_last_gram_seconds())          # last timed step, this rank's row block
    gram_alg_tflops = float(n_local) * p * (p + 1) / max(gram_kernel_s, 1e-9) / 1e12
    nb = (p + 255) // 256
    gram_exec_tflops = 3.0 * 2.0 * n_local * 65536.0 * (nb * (nb + 1) // 2) / max(gram_kernel_s, 1e-9) / 1e12
    tr_gram = ncu_traffic("gram_pair_h_kernel", n_local, p)
    tr_iter = ncu_traffic("tall_path_tri_kernel" if tri_kernel_in_use(world) else "tall_path_kernel", n, p)
    name = ("enet_tall_n%d_p%d_alpha0.5_1lambda" if enet else "lasso_tall_n%d_p%d_" + "%dlambda" % nl) % (n, p)
    line = env.base_line("admm_iters_per_sec_full_lambda_path", UNIT, total_iters / dev_s, dev_s, "strong", "f32", {
        "workload": name, "n": n, "p": p, "nlambda": nl, "standardize": True, "intercept": True, "eps_abs": 1e-5, "eps_rel": 1e-5,
        "maxit": 10000, "rho": "auto", "alpha": alpha if enet else 1.0, "lambda": lam_user,
        "sharding": "rows over %d rank(s), Gram all-reduce, iterations sharded over peer memory" % world,
        "l2": "inputs (%.1f GB) exceed L2; no flush needed" % (4.0 * n_local * p / 1e9)})
    line.update({
        "path_wall_s": dev_s / args.steps, "host_wall_s_per_step": wall_s / args.steps, "niter_path": niter_path,
        "iters_per_sec_steady": niter_path / T["iterate"], "phase_s": T,
        # the dominant kernel of the step: the Gram matrix on the tensor cores.  achieved = ALGORITHMIC flops
        # n p (p + 1) (a symmetric rank-n update) per launch / kernel time; the kernel executes 3.2x that (three fp16
        # products per element for fp32 accuracy, whole 256 x 256 tiles on the diagonal): `executed`.
        "roofline": {"kernel": "gram_pair_h_kernel (tcgen05 kind::f16 CTA-pair Gram, 3-product fp16 split)", "bound": "tensor",
                     "achieved": gram_alg_tflops, "peak": env.tensor_peak, "unit": "TFLOP/s", "frac": gram_alg_tflops / env.tensor_peak,
                     "traffic": (tr_gram["dram_bytes"] if tr_gram else None), "traffic_source": (tr_gram["source"] if tr_gram else None),
                     "peak_source": env.peak_src + ", dense bf16 sustained", "kernel_s": gram_kernel_s,
                     "algorithmic_flop": float(n_local) * p * (p + 1), "executed": gram_exec_tflops, "executed_frac": gram_exec_tflops / env.tensor_peak},
        # the per-iteration hot path named by BASELINE.json: one persistent launch for the whole lambda path.  Two
        # byte counts: SURVEY.md section 8(d)'s 4 p (p + 1) + 64 p (the reference's two triangular solves = what a full
        # K^-1 read costs) and the symmetric minimum 2 p (p + 1) + 64 p (one triangle of K^-1).
        "roofline_iteration": iteration_roofline(env, p, niter_path, T["iterate"], world, tr_iter),
        "setup_flops": {"gram_syrk_flop": gram_flops, "gram_phase_tflops": gram_flops / world / max(T["gram"], 1e-9) / 1e12,
                        "factor_flop": float(p) ** 3, "factor_tflops": float(p) ** 3 / max(T["factor"], 1e-9) / 1e12},
        "clocks": clocks, "gpu_launches": int(launches), "device": env.info["name"]})

    parity_ok = True
    # ---- parity at N > 1: every rank against one GPU's fit of the undivided problem --------------------------------
    if world > 1 and not args.no_parity:
        f = fits[-1]
        bg = np.asarray(f.beta.todense())
        # (a) all ranks hold the identical result
        t = torch.from_numpy(np.ascontiguousarray(bg)).cuda()
        t0 = t.clone()
        dist.broadcast(t0, 0)
        same = torch.tensor([1.0 if (torch.equal(t, t0) and True) else 0.0], device="cuda")
        ni = torch.from_numpy(f.niter.astype(np.int64)).cuda()
        ni0 = ni.clone()
        dist.broadcast(ni0, 0)
        if not torch.equal(ni, ni0):
            same[0] = 0.0
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        pv = {"identical_on_all_ranks": bool(same.item() == 1.0)}
        if rank == 0:
            # (b) rank 0 generates the whole design and fits it alone (communicator suspended)
            del Xd, yd
            torch.cuda.empty_cache()
            L.b200admm_release_cache()
            Xf, yf = env.synth(n, p, 0)
            L.b200admm_comm_suspend(1)
            try:
                f1 = pen(model(Xf.t(), yf)).fit()
            finally:
                L.b200admm_comm_suspend(0)
            del Xf, yf
            b1 = np.asarray(f1.beta.todense())
            cmpv = compare_paths(bg, b1, f.niter, f1.niter, None, 1e-4, 1e-4)
            cmpv["niter_n"] = cmpv.pop("niter_gpu"); cmpv["niter_1"] = cmpv.pop("niter_cpu")
            cmpv["rho_n"], cmpv["rho_1"] = f.info["rho"], f1.info["rho"]
            cmpv["ok"] = bool(cmpv["ok"] and pv["identical_on_all_ranks"] and abs(f.info["rho"] / f1.info["rho"] - 1) < 1e-4)
            pv.update(cmpv)
            parity_ok = pv["ok"]
            Xd, yd = env.synth(n_local, p, row0)
        env.barrier()
        line["parity_vs_n1"] = pv

    # ---- parity at N = 1: the CPU oracle on the GPU's own Gram matrix -----------------------------------------------
    cpu_from_parity = None
    if world == 1 and not args.no_parity:
        from oracle import pyoracle as O
        cores = host_threads()
        bt = O.use_openblas(cores)
        O.omp_threads(cores)
        with K.capture(p) as cap:
            fpar = fit_device()
        bg = np.asarray(fpar.beta.todense())
        par = {"checker": "CPU oracle (oracle/admm_oracle.cpp) run on the GPU's own X'X, X'y and lambda grid; rho from the oracle's own Lanczos"}
        par["gram_spot_check"] = gram_spot_check(env, Xd, n_local, cap)
        bc, o, t_iter_cpu, t_setup_cpu = oracle_on_captured_gram(O, cap, fpar, n, enet=enet, alpha=alpha)
        # bounds: see compare_paths (1e-4 where both runs stop at the same iteration; the stopping rule's own slack elsewhere)
        par.update(compare_paths(bg, bc, fpar.niter, o["niter"], None, 1e-4, 1e-4))
        par["rho_gpu"], par["rho_cpu"] = fpar.info["rho"], float(o["rho"])
        par["eig_gpu"], par["eig_cpu"] = fpar.info["eig"], float(o["eig"])
        par["same_as_timed_fit"] = bool(np.array_equal(bg, np.asarray(fits[-1].beta.todense())))
        par["ok"] = bool(par["ok"] and par["gram_spot_check"]["max_err_over_n"] < 2e-5 and abs(par["rho_gpu"] / par["rho_cpu"] - 1) < 1e-4)
        # full-design CPU fit of the reference arm, if it ran on this box (bit-identical X by construction)
        rf = ref_fit_file(args)
        if os.path.exists(rf):
            try:
                z = np.load(rf)
                if z["beta"].shape == bg.shape:
                    # both sides accumulate a 1e6-term float32 Gram matrix in their own order (the GPU's is within 3e-9 n of
                    # float64, the CPU's blocked SYRK is not), so this comparison carries that difference too: 5e-4 where
                    # the iteration counts agree, the stopping rule's slack elsewhere
                    pr = compare_paths(bg, z["beta"], fpar.niter, z["niter"], None, 5e-4, 1e-4)
                    pr["source"] = "bench.py --impl reference on this box: oracle fit of the full n x p design from the bit-identical CPU generator"
                    par["vs_reference_arm_full_fit"] = pr
                    par["ok"] = bool(par["ok"] and pr["ok"])
            except Exception as ex:
                par["vs_reference_arm_full_fit"] = {"error": repr(ex)[:200]}
        line["parity"] = par
        parity_ok = par["ok"]
        cpu_from_parity = (O, bt, cores, int(o["niter"].sum()), t_iter_cpu, t_setup_cpu)

    # ---- end to end: host buffers in, host results out, every step ----------------------------------------------------
    if not args.no_e2e:
        numa = bind_near_gpu(env.local_rank)
        try:
            Xh = torch.empty((p, n_local), dtype=torch.float32, pin_memory=True)
            yh = torch.empty(n_local, dtype=torch.float32, pin_memory=True)
            Xh.copy_(Xd); yh.copy_(yd)
            torch.cuda.synchronize()
            xh_np, yh_np = Xh.numpy().T, yh.numpy()

            def fit_host():
                return pen(model(xh_np, yh_np)).fit()
            fit_host()
            fh, e_dev, e_wall, _ = env.timed(fit_host, args.e2e_steps)
            it = sum(int(f.niter.sum()) for f in fh)
            h2d = int(4 * n_local * p + 4 * n_local)
            gsec = float(np.mean([f.info["timing"]["gram"] for f in fh]))
            rates = [h2d / max(gsec, 1e-9) / 1e9]
            if world > 1:
                rt = torch.zeros(world, dtype=torch.float64, device="cuda")
                rt[rank] = rates[0]
                dist.all_reduce(rt)
                rates = [float(v) for v in rt]
            line["e2e"] = {"value": it / e_wall, "unit": UNIT, "h2d_bytes_per_step": h2d,
                           "d2h_bytes_per_step": int(4 * nl * p + 4 * nl + 3 * 4 * p + 8),
                           "path_wall_s": e_wall / args.e2e_steps, "steps": args.e2e_steps, "device_s_per_step": e_dev / args.e2e_steps,
                           "phase_s": {k: float(np.mean([f.info["timing"][k] for f in fh])) for k in fh[-1].info["timing"]},
                           "phase_s_per_step": [{k: round(float(v), 4) for k, v in f.info["timing"].items()} for f in fh],
                           "h2d_gbs_per_rank": [round(r, 2) for r in rates],
                           "h2d_note": "this rank's bytes / its copy-overlapped ingest+DataStd+Gram phase (copy-bound at N = 1)",
                           "host_memory": numa, "input": "float32 column-major, pinned host memory",
                           # bit-identical on one GPU (the pipelined ingest adds the same K-slices in the same order); on N > 1 the
                           # per-panel all-reduces sum the ranks' partial Gram matrices panel by panel: same values up to rounding
                           "same_result_as_device_input": bool(np.array_equal(np.asarray(fh[-1].beta.todense()), np.asarray(fits[-1].beta.todense()))),
                           "max_abs_dbeta_vs_device_input": float(np.abs(np.asarray(fh[-1].beta.todense()) - np.asarray(fits[-1].beta.todense())).max()),
                           "niter_vs_device_input": [int(fh[-1].niter.sum()), int(fits[-1].niter.sum())]}
            del Xh, yh, xh_np, yh_np
            # R's own layout: float64 host, narrowed on the device (Lasso.cpp:45-50) -- at a size an R session can hold
            if world == 1 and not enet:
                n64 = min(n_local, 200_000)
                X64 = torch.empty((p, n64), dtype=torch.float64, pin_memory=True)
                X64.copy_(Xd[:, :n64])
                y64 = yd[:n64].double().cpu().numpy()
                x64_np = X64.numpy().T

                def fit_host64():
                    return A.admm_lasso(x64_np, y64).penalty(nlambda=nl).fit()
                fit_host64()
                f64, d64, w64, _ = env.timed(fit_host64, 2)
                line["e2e_f64_host"] = {"value": sum(int(f.niter.sum()) for f in f64) / w64, "unit": UNIT, "n": n64, "p": p,
                                        "h2d_bytes_per_step": int(8 * n64 * p + 8 * n64), "path_wall_s": w64 / 2,
                                        "phase_s": {k: float(np.mean([f.info["timing"][k] for f in f64])) for k in f64[-1].info["timing"]},
                                        "input": "float64 column-major (R's REALSXP layout), pinned host memory; narrowed to float32 on the device"}
                del X64, x64_np
        except Exception as ex:  # pinned allocation can fail on small hosts: say so, do not fake a number
            line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)[:300]}

    # ---- CPU baseline (N = 1): measured iteration phase at full size + sampled DataStd / Gram -------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        del Xd
        torch.cuda.empty_cache()
        try:
            if cpu_from_parity is None:
                raise RuntimeError("needs the parity run (its oracle path is the measured iteration phase)")
            O, bt, cores, nit_cpu, t_iter_cpu, t_setup_cpu = cpu_from_parity
            ns = min(args.cpu_rows, n)
            t_std, t_gram = cpu_std_gram_sample(O, args, ns)
            full_wall = (t_std + t_gram) * (n / ns) + t_setup_cpu + t_iter_cpu
            line["cpu_baseline"] = {
                "value": nit_cpu / full_wall, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": ("oracle (restated reference, OpenBLAS %d threads): the whole %d-lambda iteration phase at full p on the GPU run's own "
                           "Gram matrix (%d iterations, %.2f ms/iter) and Lanczos + Cholesky are measured; DataStd + Gram timed on %d of %d "
                           "rows of the same design and scaled linearly in n.  `bench.py --impl reference` measures the whole fit."
                           % (bt, nl, nit_cpu, t_iter_cpu / max(nit_cpu, 1) * 1e3, ns, n)),
                "path_wall_s_extrapolated": full_wall, "iters_per_s_steady": nit_cpu / t_iter_cpu,
                "measured_s": {"standardize_sample": t_std, "gram_sample": t_gram, "lanczos_cholesky": t_setup_cpu, "iterations": t_iter_cpu}}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "error": repr(ex)[:300]}

    if rank == 0:
        print(json.dumps(line), flush=True)
    env.finish()
    if world > 1:
        ok = parity_ok          # rank 0 decides; the others exit 0
    else:
        ok = parity_ok
    if rank == 0 and not ok:
        sys.stderr.write("bench.py: PARITY MISS -- see the `parity` / `parity_vs_n1` object of the line above\n")
        sys.exit(3)


# ---------------------------------------------------------------------------------------------------------------
# --config wide : lasso n <= p (C3: n = 1e4, p = 1e6, 100 lambdas)
# ---------------------------------------------------------------------------------------------------------------
def kkt_lasso(env, Xd, yd, beta, lam):
    """Size-independent optimality check of a lasso solution on standardised data (torch on the GPU as checker):
    gradient g_j = x_j'(y - b0 - X b) / (n sd_j); |g_j| <= lambda and g_j = lambda sign(b_j) on the support.
    Returns the worst violations relative to lambda."""
    torch = env.torch
    p, n = Xd.shape
    b = torch.from_numpy(beta[1:]).cuda().float()
    nz = b.nonzero().flatten()
    r = yd.double() - float(beta[0])
    if nz.numel():
        r = r - (Xd[nz].double().t() @ b[nz].double())
    r32 = r.float()
    g = torch.empty(p, dtype=torch.float64, device="cuda")
    sd = torch.empty(p, dtype=torch.float64, device="cuda")
    step = max(1, (1 << 28) // n)
    for j0 in range(0, p, step):
        blk = Xd[j0:j0 + step]
        g[j0:j0 + step] = (blk @ r32).double()
        m = blk.mean(dim=1, keepdim=True)
        sd[j0:j0 + step] = ((blk - m).double().pow(2).sum(dim=1) / n).sqrt()
    g = g / n / sd
    viol = float((g.abs().max() / lam - 1.0).clamp(min=0))
    on = 0.0
    if nz.numel():
        on = float(((g[nz] - lam * torch.sign(b[nz]).double()).abs() / lam).max())
    return {"max_excess_over_lambda_rel": viol, "max_support_gradient_err_rel": on, "support": int(nz.numel())}


def run_wide(args):
    env = Env(args)
    torch, A, K, L = env.torch, env.A, env.K, env.L
    if env.rank != 0:
        env.finish()
        return
    n, p, nl = args.n, args.p, args.nlambda
    Xd, yd = env.synth(n, p, 0)

    def fit_device():
        return A.admm_lasso(Xd.t(), yd).penalty(nlambda=nl).opts(maxit=args.maxit).fit()
    for _ in range(min(args.warmup, 1)):
        fit_device()
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    steps = min(args.steps, 2)
    args.steps = steps
    fits, dev_s, wall_s, launches = env.timed(fit_device, steps)
    clocks = sampler.stop()
    work = np.zeros(4)
    L.b200admm_last_work(work.ctypes.data)
    f = fits[-1]
    T = f.info["timing"]
    nit = int(f.niter.sum())
    achieved = work[0] / T["iterate"] / 1e9
    line = env.base_line("admm_iters_per_sec_full_lambda_path", UNIT, sum(int(q.niter.sum()) for q in fits) / dev_s, dev_s, "weak", "f32", {
        "workload": "lasso_wide_n%d_p%d_%dlambda" % (n, p, nl), "n": n, "p": p, "nlambda": nl, "standardize": True, "intercept": True,
        "eps_abs": 1e-5, "eps_rel": 1e-5, "maxit": args.maxit, "rho": "auto", "lambda_min_ratio": 0.01,
        "l2": "X (%.1f GB) exceeds L2; no flush needed" % (4.0 * n * p / 1e9)})
    line.update({"path_wall_s": dev_s / steps, "niter_path": nit, "phase_s": T, "us_per_iteration": T["iterate"] / max(nit, 1) * 1e6,
                 "regular_steps": int(work[1]), "active_set_steps": int(work[2]), "gamma": f.info["eig"], "rho_final": f.info["rho"],
                 "support_last_lambda": int(f.beta[:, -1].nnz) - 1,
                 "roofline": {"kernel": "wide iteration kernels (wide_screen over the fp16 copy + gemv_t_list on regular steps, wide_active / wide_ax column gathers otherwise)",
                              "bound": "hbm", "achieved": achieved, "peak": env.hbm_peak, "unit": "GB/s", "frac": achieved / env.hbm_peak,
                              "traffic": None, "algorithmic_bytes": float(work[0]),
                              "note": "sum over iterations of [2 n p (fp16 screen) + 4 n candidates] (screened regular step; 4 n p unscreened) or "
                                      "4 n nnz_k (active set), + 4 n nnz_{k+1} + 48 n, / iterate seconds",
                              "screened": os.environ.get("B200ADMM_WIDE_SCREEN", "1") != "0",
                              "mean_columns_evaluated_exactly_per_regular_step": float(work[3]),
                              "peak_source": env.peak_src},
                 "clocks": clocks, "gpu_launches": int(launches), "device": env.info["name"]})
    ok = True
    if not args.no_parity:
        # full-size parity through the optimality conditions (the CPU oracle needs ~10 minutes for this path at this
        # size; tests/test_gpu_models.py compares with it at n = 2000 x p = 20000): last lambda and one mid-path lambda
        B = np.asarray(f.beta.todense())
        par = {"checker": "lasso optimality conditions on the standardised design, torch float64 on the GPU", "lambdas": {}}
        for k in (nl // 2, nl - 1):
            kk = kkt_lasso(env, Xd, yd, B[:, k], float(f.lambda_[k]))
            par["lambdas"]["%d" % k] = kk
            # the reference stops at eps = 1e-5 relative residuals: its own README run is 2e-3 from glmnet at p > n
            ok = ok and kk["max_excess_over_lambda_rel"] < 0.05 and kk["max_support_gradient_err_rel"] < 0.05
        par["ok"] = bool(ok)
        line["parity"] = par
    if not args.no_e2e:
        try:
            numa = bind_near_gpu(env.local_rank)
            Xh = torch.empty((p, n), dtype=torch.float32, pin_memory=True)
            Xh.copy_(Xd)
            yh = yd.cpu().numpy()
            xh = Xh.numpy().T
            del Xd
            torch.cuda.empty_cache()

            def fit_host():
                return A.admm_lasso(xh, yh).penalty(nlambda=nl).opts(maxit=args.maxit).fit()
            fit_host()
            fh, e_dev, e_wall, _ = env.timed(fit_host, 1)
            line["e2e"] = {"value": int(fh[0].niter.sum()) / e_wall, "unit": UNIT, "h2d_bytes_per_step": int(4 * n * p + 4 * n),
                           "d2h_bytes_per_step": int(8 * fh[0].beta.nnz + 8 * nl), "path_wall_s": e_wall, "steps": 1,
                           "phase_s": fh[0].info["timing"], "host_memory": numa, "input": "float32 column-major, pinned host memory"}
            del Xh, xh
        except Exception as ex:
            line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)[:300]}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_wide_sample(args)
    print(json.dumps(line), flush=True)
    env.finish()
    if not ok:
        sys.exit(3)


def cpu_wide_sample(args, p_s=50_000, nl_s=100):
    """The oracle's wide path on a narrower design of the same n (p_s columns of the same generator, all lambdas):
    the per-iteration cost is proportional to the columns touched, so seconds are reported per algorithmic byte and
    scaled to the full run's byte count by the caller's reader -- stated, not hidden."""
    try:
        from oracle import pyoracle as O
        cores = host_threads()
        bt = O.use_openblas(cores)
        O.omp_threads(cores)
        X, y = O.synth_f32(args.n, p_s, row0=0, seed=args.seed, nsig=min(100, p_s))
        t0 = time.perf_counter()
        o = O.lasso_path(X.astype(np.float64), y.astype(np.float64), nlambda=nl_s, maxit=args.maxit)
        wall = time.perf_counter() - t0
        nit = int(o["niter"].sum())
        return {"value": nit / wall, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "oracle (restated reference, OpenBLAS %d threads) on n = %d x p = %d (the first %d columns' worth of the same generator), "
                          "%d lambdas: %d iterations in %.1f s.  The full configuration has %dx the columns: setup (XX' Gram) and the regular "
                          "steps scale with p, the active-set steps with the support" % (bt, args.n, p_s, p_s, nl_s, nit, wall, args.p // p_s),
                "path_wall_s": wall, "niter": nit}
    except Exception as ex:
        return {"value": None, "error": repr(ex)[:300]}


# ---------------------------------------------------------------------------------------------------------------
# --config lad / bp : float64 solvers (C4: LAD n = 5e5 x p = 5e3; BP n = 5e3 x p = 5e5)
# ---------------------------------------------------------------------------------------------------------------
def run_lad_bp(args, which):
    env = Env(args)
    torch, A, K, L = env.torch, env.A, env.K, env.L
    if env.rank != 0:
        env.finish()
        return
    n, p = args.n, args.p
    if which == "lad":
        X32, y32 = env.synth(n, p, 0)
        Xd, yd = X32.double(), y32.double()
        del X32
    else:
        # noiseless sparse recovery (README.md:371-377 pattern): 500 signals among p columns, y = A beta*
        X32, _ = env.synth(n, p, 0, nsig=0, noise=0.0, sd=1.0)
        Xd = X32.double()
        del X32
        g = torch.Generator(device="cuda").manual_seed(args.seed)
        bt = torch.zeros(p, dtype=torch.float64, device="cuda")
        idx = torch.randperm(p, device="cuda", generator=g)[:500]
        bt[idx] = torch.rand(500, dtype=torch.float64, device="cuda", generator=g)
        yd = Xd.t() @ bt
    torch.cuda.empty_cache()

    def fit_device(maxit=args.maxit):
        m = A.admm_lad(Xd.t(), yd) if which == "lad" else A.admm_bp(Xd.t(), yd)
        return m.opts(maxit=maxit).fit()
    fit_device(maxit=5)
    steps = 1
    args.steps = steps
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    fits, dev_s, wall_s, launches = env.timed(fit_device, steps)
    clocks = sampler.stop()
    f = fits[-1]
    T = f.info["timing"]
    nit = int(f.niter)
    bpi = 2.0 * n * p * 8 + (p * p * 8 if which == "lad" else 0) + 16.0 * (n if which == "lad" else p) * 8
    achieved = bpi * min(nit, args.maxit) / T["iterate"] / 1e9
    line = env.base_line("admm_iters_per_sec_full_fit", UNIT_FIT, nit / dev_s, dev_s, "weak", "f64", {
        "workload": "%s_n%d_p%d" % (which, n, p), "n": n, "p": p, "eps_abs": 1e-4, "eps_rel": 1e-4, "maxit": args.maxit, "rho": 1.0,
        "l2": "X (%.1f GB float64) exceeds L2; no flush needed" % (8.0 * n * p / 1e9)})
    line.update({"fit_wall_s": dev_s, "niter": nit, "phase_s": T, "ms_per_iteration": T["iterate"] / max(nit, 1) * 1e3, "rho_final": f.info["rho"],
                 "roofline": {"kernel": "gemv_t / gemv_n float64 over %s (two passes per iteration)" % ("X" if which == "lad" else "M = L^-1 A"),
                              "bound": "hbm", "achieved": achieved, "peak": env.hbm_peak, "unit": "GB/s", "frac": achieved / env.hbm_peak,
                              "traffic": None, "bytes_per_iteration": bpi, "peak_source": env.peak_src},
                 "setup_flops": {"gram_tflops_f64": float(n) * p * (p + 1) / max(T["gram"], 1e-9) / 1e12 if which == "lad"
                                 else float(p) * n * (n + 1) / max(T["gram"], 1e-9) / 1e12},
                 "clocks": clocks, "gpu_launches": int(launches), "device": env.info["name"]})
    ok = True
    if not args.no_parity or not args.no_cpu:
        # full-size parity AND the CPU baseline in one oracle run: the first 20 iterations of the reference algorithm on
        # the same matrix (copied to the host), trace (eps, residuals, rho) and coefficients compared
        try:
            from oracle import pyoracle as O
            cores = host_threads()
            btn = O.use_openblas(cores)
            O.omp_threads(cores)
            nit_s, extra = 20, 6
            with K.trace(which=0, cap=nit_s + extra + 5) as tr:
                fs = fit_device(maxit=nit_s + extra)
            xh = Xd.cpu().numpy().T                                  # (n, p), Fortran-ordered view of the (p, n) copy
            yh = yd.cpu().numpy()
            t0 = time.perf_counter()
            o = (O.lad(xh, yh, maxit=nit_s, trace_cap=nit_s + 5) if which == "lad" else O.bp(xh, yh, maxit=nit_s, trace_cap=nit_s + 5))
            wall = time.perf_counter() - t0
            # Both solvers start with a run of iterations in which z does not move; the x-update then returns the same point
            # and the restart rule `c < 0.999 c_old` (src/FADMMBase.h:243) compares a number with 0.999 * (itself / 0.999) --
            # decided by the last bit of |r|^2, i.e. by the summation order of a norm.  A run that takes the restart branch
            # repeats the iteration (a "stutter" row: dual residual exactly 0, primal residual unchanged) and is then the
            # other run shifted by one iteration (measured at n = 5e5 x p = 5e3: all 15 later rows agree to 1e-14 after the
            # shift, profiles/r2i_lad_bp_trace_forensics.log).  So: rows before the first such decision must agree to 1e-9,
            # the traces must align modulo stutter rows, and (LAD) the coefficients at the aligned iteration must agree.
            # BP's rho balancing depends on the iteration INDEX (i > 5, src/FADMMBase.h:250), so its shifted runs part for
            # good; there the converged GPU solution is certified instead (feasible within what the stopping rule allows,
            # l1 norm not above the planted signal's).
            tc = o["trace"][:min(int(o["niter"]), nit_s)]
            al = align_traces(tr.rows, tc, 1e-9)
            par = {"checker": "CPU oracle on the same float64 matrix, first %d iterations: per-iteration (eps, residuals, rho) aligned modulo "
                              "restart-rule stutter rows" % nit_s,
                   "trace_rows_cpu": int(len(tc)), "rows_equal_before_first_knife_edge": al["prefix"], "rows_aligned": al["matched"],
                   "gpu_stutter_rows": al["skipped_a"], "cpu_stutter_rows": al["skipped_b"], "trace_max_rel_diff_aligned": al["max_rel"],
                   "aligned_to_the_end": al["ok"]}
            if which == "lad":
                par["ok"] = False
                if al["ok"] and al["last_a"] >= 0:
                    f2 = fit_device(maxit=al["last_a"] + 1)
                    db = float(np.abs(f2.beta - o["beta"]).max())
                    par.update({"gpu_iterations_at_cpu_iteration_%d" % len(tc): al["last_a"] + 1, "max_abs_dbeta": db,
                                "beta_inf": float(np.abs(o["beta"]).max())})
                    par["ok"] = bool(al["prefix"] >= 3 and al["max_rel"] < 1e-9 and db < 1e-7 * max(1.0, par["beta_inf"]))
            else:
                bfull = torch.from_numpy(np.asarray(f.beta.todense())[:, 0]).cuda()
                res2 = float((Xd.t() @ bfull - yd).norm())
                eps_pri = (p ** 0.5) * 1e-4 + 1e-4 * float(bfull.norm())
                bound = 1.1 * (n ** 0.5 + p ** 0.5) * eps_pri               # |A (z - x)| <= |A|_2 |z - x|, |A|_2 ~ sqrt(n) + sqrt(p)
                l1, l1_true = float(bfull.abs().sum()), float(bt.abs().sum())
                par.update({"converged_fit": {"residual_2norm": res2, "residual_bound_from_stopping_rule": bound,
                                              "residual_rel_inf": float((Xd.t() @ bfull - yd).abs().max() / yd.abs().max()),
                                              "l1_norm": l1, "l1_norm_planted_signal": l1_true,
                                              "max_abs_err_vs_planted": float((bfull - bt).abs().max()), "niter": nit}})
                par["ok"] = bool(al["prefix"] >= 2 and res2 <= bound and l1 <= l1_true * (1 + 1e-3) and nit <= args.maxit)
            ok = par["ok"]
            line["parity"] = par
            line["cpu_baseline"] = {"value": nit_s / wall, "unit": UNIT_FIT, "cores": cores, "kind": "port",
                                    "sample": "oracle (restated reference, OpenBLAS %d threads) on the full n = %d x p = %d float64 matrix: setup (Gram, Cholesky%s) "
                                              "+ the first %d iterations, %.1f s" % (btn, n, p, ", M = L^-1 A" if which == "bp" else "", nit_s, wall),
                                    "wall_s": wall, "iterations": nit_s}
        except Exception as ex:
            line["parity"] = {"ok": False, "error": repr(ex)[:300]}
            ok = False
    if not args.no_e2e:
        try:
            numa = bind_near_gpu(env.local_rank)
            Xh = torch.empty((p, n), dtype=torch.float64, pin_memory=True)
            Xh.copy_(Xd)
            yh = yd.cpu().numpy()
            xh = Xh.numpy().T
            del Xd
            torch.cuda.empty_cache()

            def fit_host():
                m = A.admm_lad(xh, yh) if which == "lad" else A.admm_bp(xh, yh)
                return m.opts(maxit=args.maxit).fit()
            fh, e_dev, e_wall, _ = env.timed(fit_host, 1)
            line["e2e"] = {"value": int(fh[0].niter) / e_wall, "unit": UNIT_FIT, "h2d_bytes_per_step": int(8 * n * p + 8 * n),
                           "d2h_bytes_per_step": int(8 * (p + 1)), "fit_wall_s": e_wall, "steps": 1, "phase_s": fh[0].info["timing"],
                           "host_memory": numa, "input": "float64 column-major (R's layout), pinned host memory"}
        except Exception as ex:
            line["e2e"] = {"value": None, "unit": UNIT_FIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)[:300]}
    print(json.dumps(line), flush=True)
    env.finish()
    if not ok:
        sys.exit(3)


# ---------------------------------------------------------------------------------------------------------------
# --config consensus : the reference's $parallel(N) (C5: n = 1e6 x p = 8e4 over 2 / 4 / 8 GPUs)
# ---------------------------------------------------------------------------------------------------------------
def run_consensus(args):
    env = Env(args)
    torch, dist, A, K, L = env.torch, env.dist, env.A, env.K, env.L
    rank, world = env.rank, env.world
    from admm_b200 import dist as D
    n, p = args.n, args.p
    nblocks = world if world > 1 else 2
    if world > 1:
        r0, nr = D.row_block(n, world, rank)
    else:
        r0, nr = 0, n
    Xd, yd = env.synth(nr, p, r0)
    f0 = A.admm_lasso(Xd.t(), yd).penalty(nlambda=2).parallel(nblocks).opts(maxit=1).fit()
    lam = [0.1 * float(f0.lambda_[0])]

    def fit_device():
        return A.admm_lasso(Xd.t(), yd).penalty(lam).parallel(nblocks).opts(maxit=args.maxit).fit()
    steps = 1
    args.steps = steps
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    fits, dev_s, wall_s, launches = env.timed(fit_device, steps)
    clocks = sampler.stop()
    f = fits[-1]
    T = f.info["timing"]
    it = int(f.niter.sum())
    tt = torch.tensor([T["gram"], T["iterate"], T["standardize"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_setup, t_iter, t_std = (float(v) for v in tt)
    bpi = tall_bytes_per_iter(p) * (1 if world > 1 else nblocks)
    achieved = bpi * min(it, args.maxit) / max(t_iter, 1e-9) / 1e9
    line = env.base_line("admm_iters_per_sec_full_lambda_path", UNIT, it / dev_s, dev_s, "strong", "f32", {
        "workload": "consensus_lasso_rowsplit_n%d_p%d_%dblocks_1lambda" % (n, p, nblocks), "n": n, "p": p, "blocks": nblocks,
        "lambda": lam[0], "eps_abs": 1e-5, "eps_rel": 1e-5, "maxit": args.maxit, "rho": "lambda / N (reference default)",
        "sharding": "one row block per GPU, one all-reduce of p + 3 floats per iteration" if world > 1 else "%d blocks on one GPU" % nblocks,
        "l2": "K_i^-1 (%.1f GB) exceeds L2; no flush needed" % (4.0 * p * p / 1e9)})
    line.update({"fit_wall_s": dev_s, "niter": it, "converged": bool(it <= args.maxit), "phase_s": T,
                 "decomposition_s": {"standardize": t_std, "block_gram_and_inverse": t_setup, "iterations": t_iter,
                                     "note": "setup scales with 1/N (rows per block); per-iteration work per GPU does not (each block solves a full p x p system) and the iteration count grows with N (SURVEY.md D8)"},
                 "ms_per_iteration": t_iter / max(min(it, args.maxit), 1) * 1e3,
                 "roofline": {"kernel": "cons_x_kernel (K_i^-1 product per block)", "bound": "hbm", "achieved": achieved, "peak": env.hbm_peak,
                              "unit": "GB/s", "frac": achieved / env.hbm_peak, "traffic": None, "bytes_per_iteration_per_gpu": bpi,
                              "peak_source": env.peak_src, "allreduce_payload_bytes": 4 * (p + 3)},
                 "clocks": clocks, "gpu_launches": int(launches), "device": env.info["name"]})
    ok = True
    if not args.no_parity:
        # all ranks hold the identical result; optimality conditions of the lasso at the returned point (distributed)
        bg = np.asarray(f.beta.todense())[:, 0]
        same = torch.tensor([1.0], device="cuda")
        if world > 1:
            t = torch.from_numpy(bg.copy()).cuda()
            t0 = t.clone()
            dist.broadcast(t0, 0)
            if not torch.equal(t, t0):
                same[0] = 0.0
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
        b = torch.from_numpy(bg[1:]).cuda().float()
        nz = b.nonzero().flatten()
        r = yd.double() - float(bg[0])
        if nz.numel():
            r = r - Xd[nz].double().t() @ b[nz].double()
        r32 = r.float()
        stats = torch.zeros((3, p), dtype=torch.float64, device="cuda")       # X'r, column sums, column sums of squares
        step = max(1, (1 << 28) // nr)
        for j0 in range(0, p, step):
            blk = Xd[j0:j0 + step]
            stats[0, j0:j0 + step] = (blk @ r32).double()
            stats[1, j0:j0 + step] = blk.double().sum(dim=1)
            stats[2, j0:j0 + step] = blk.double().pow(2).sum(dim=1)
        if world > 1:
            dist.all_reduce(stats)
        mean = stats[1] / n
        sd = (stats[2] / n - mean * mean).clamp(min=1e-30).sqrt()
        g = stats[0] / n / sd
        lam0 = lam[0]
        viol = float((g.abs().max() / lam0 - 1.0).clamp(min=0))
        on = float(((g[nz] - lam0 * torch.sign(b[nz]).double()).abs() / lam0).max()) if nz.numel() else 0.0
        par = {"checker": "lasso optimality conditions on the standardised design (torch float64, all-reduced over the row blocks)",
               "identical_on_all_ranks": bool(same.item() == 1.0), "max_excess_over_lambda_rel": viol, "max_support_gradient_err_rel": on,
               "support": int(nz.numel())}
        par["ok"] = bool(par["identical_on_all_ranks"] and viol < 0.05 and on < 0.05 and it <= args.maxit)
        ok = par["ok"]
        line["parity"] = par
    if rank == 0:
        print(json.dumps(line), flush=True)
    env.finish()
    if rank == 0 and not ok:
        sys.exit(3)


# ---------------------------------------------------------------------------------------------------------------
# --impl reference : the oracle on the host cores
# ---------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as O
    cores = host_threads()
    bt = O.use_openblas(cores)
    O.omp_threads(cores)
    t_all = time.perf_counter()
    n, p, nl = args.n, args.p, args.nlambda
    cfg = {"workload": None, "n": n, "p": p}
    if args.config in ("tall", "enet"):
        enet = args.config == "enet"
        name = ("enet_tall_n%d_p%d_alpha0.5_1lambda" if enet else "lasso_tall_n%d_p%d_" + "%dlambda" % nl) % (n, p)
        cfg.update({"workload": name, "nlambda": 1 if enet else nl})
        if args.ref_sample:
            ns = min(args.cpu_rows, n)
            t_std, t_gram = cpu_std_gram_sample(O, args, ns)
            r = O.tall_fit_synth(ns * 8, p, seed=args.seed, nsig=min(100, p), nlambda=nl, enet=enet, alpha=0.5, lambda_frac=0.1 if enet else 0.0)
            nit = int(r["niter"].sum())
            wall = (t_std + t_gram) * (n / ns) + r["times"]["lanczos_cholesky"] + r["times"]["iterations"]
            sample = "BOUNDED SAMPLE (--ref-sample): DataStd + Gram on %d rows scaled linearly in n; path measured on a %d-row design" % (ns, ns * 8)
            meas = {"standardize_sample": t_std, "gram_sample": t_gram, **r["times"]}
            steps_measured, extrapolated = 0, True
        else:
            r = O.tall_fit_synth(n, p, seed=args.seed, nsig=min(100, p), nlambda=nl, enet=enet, alpha=0.5, lambda_frac=0.1 if enet else 0.0,
                                 chunk_rows=32768)
            nit = int(r["niter"].sum())
            t = r["times"]
            wall = t["standardize"] + t["gram"] + t["lanczos_cholesky"] + t["iterations"]
            sample = ("oracle (restated reference, OpenBLAS %d threads) on the FULL design, nothing extrapolated: n = %d rows streamed in 32768-row chunks "
                      "from the CPU generator (bit-identical to the GPU arm's X; %.1f s of generation NOT counted -- X is the caller's input), "
                      "DataStd %.1f s, X'y + SYRK Gram %.1f s, Lanczos + Cholesky %.2f s, %d iterations over %d lambda(s) %.1f s"
                      % (bt, n, t["generate"], t["standardize"], t["gram"], t["lanczos_cholesky"], nit, len(r["niter"]), t["iterations"]))
            meas = t
            steps_measured, extrapolated = 1, False
            try:
                np.savez(ref_fit_file(args), beta=r["beta"], niter=r["niter"], lambda_=r["lambda_"], rho=r["rho"])
            except Exception:
                pass
        value, unit, metric = nit / wall, UNIT, "admm_iters_per_sec_full_lambda_path"
        extra = {"niter_path": nit, "rho": r["rho"], "iters_per_s_steady": nit / max(r["times"]["iterations"], 1e-9)}
    elif args.config == "wide":
        cb = cpu_wide_sample(args)
        value, unit, metric, wall = cb["value"], UNIT, "admm_iters_per_sec_full_lambda_path", cb.get("path_wall_s", 0.0)
        cfg.update({"workload": "lasso_wide_n%d_p%d_%dlambda" % (n, p, nl), "nlambda": nl})
        sample, meas, steps_measured, extrapolated, extra = cb.get("sample"), {"path_wall_s": wall}, 0, True, {"niter_path": cb.get("niter")}
    else:
        print(json.dumps({"impl": "reference", "unavailable": "the CPU arm of --config %s runs inside the GPU arm (cpu_baseline: the oracle on the same matrix, "
                          "which this arm cannot regenerate without the device)" % args.config}), flush=True)
        return
    cb = {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample, "measured_s": meas,
          "extrapolated": extrapolated, "steps_measured": steps_measured}
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "steps_measured": steps_measured, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": cb,
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}
    line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "tall":
        run_tall(args, enet=False)
    elif args.config == "enet":
        run_tall(args, enet=True)
    elif args.config == "wide":
        run_wide(args)
    elif args.config in ("lad", "bp"):
        run_lad_bp(args, args.config)
    else:
        run_consensus(args)


if __name__ == "__main__":
    main()
