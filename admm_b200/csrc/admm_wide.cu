// admm_wide.cu -- lasso / elastic net for n <= p: the reference's linearised ADMM with
// active-set updates (float32 vectors, double scalars).
//
// Reference being replaced (all in /root/reference/src):
//   ADMMBase::solve / update_x / update_z / update_y / update_rho   ADMMBase.h:73-109,158-216
//   ADMMLassoWide ctor / init / init_warm                           ADMMLassoWide.h:189-251
//   ADMMLassoWide::next_x / active_set_update / is_regular_update   ADMMLassoWide.h:86-155
//   ADMMLassoWide::next_z / next_residual / compute_eps_* / resid   ADMMLassoWide.h:156-186
//   ADMMEnetWide::next_x / active_set_update / enet                 ADMMEnet.h:62-154
//   the wide branch of admm_lasso() / admm_enet()                   Lasso.cpp:75-76,112-123
//
// x (length p) is kept dense together with its sorted support list; the three data passes of
// an iteration are HBM streams over columns of X (column-major, so a column is contiguous):
//   regular step  (iterations 4^k - 1):  vec = X' tmp over ALL p columns (gemv_t), prox, support rebuild
//   active step   (all others)        :  one warp per support column: x_j - X_j' tmp, prox, prune
//   z step                            :  Ax = sum over support columns (row-parallel, fixed chunk order),
//                                        fused with z = (y_dat + u + rho Ax)/(-1-rho), r = Ax + z, u += rho r
//                                        and the five squared norms the stopping rule needs.
// The loop is driven from the host (the support size is data dependent) with ONE small device->host read per
// iteration (the five norms + the new support size); the z step sizes its launches from the previous support
// size (an upper bound during active-set steps) and reads the exact count on the device.  The solution of each
// lambda leaves the device as (index, value) pairs of its support, never as a dense p-vector.
// Compiled with --fmad=false (unfused, reference order).
#include "solvers.h"
#include "kernels.h"
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace b200 {

void finish_lasso_path(const std::vector<float>& z_all, int nl, i64 p, int flag, const std::vector<float>& meanX,
                       const std::vector<float>& scaleX, float meanY, float scaleY, b200admm_path* out);

namespace {

constexpr int WT = 256;

struct WideProx {
    int enet;
    double pen_d;      // lambda / (rho * gamma) in double (regular step, lasso)
    float pen_f;       // the same rounded to float (active step; enet parameters)
    float thresh, denom;
};

// device-side loop control of runs of active-set steps (see wide_control_kernel)
struct WideCtl {
    double rho, sAx2, sz2, sy2;                          // rho and the squared norms of the current Ax, z, y
    double eps_primal, resid_primal, eps_dual, resid_dual;   // of the last executed iteration
    double work_bytes;                                   // algorithmic bytes of the executed iterations
    float frho, pen_f, thresh, denom, den;               // parameters of the next active-set iteration
    int stop, iters, nnz;                                // converged flag, iterations executed, support size
};
struct WideCtlConst {                                    // per-lambda constants of the control kernel
    double eps_abs, eps_rel, sqrt_n, sqrt_p, lambda;
    float gamma, sqrt_sprad, alpha_f;
    int enet, i_start;
    long long n;
};

__device__ __forceinline__ float prox_regular(float v, const WideProx& q)
{
    if (!q.enet) {
        if ((double)v > q.pen_d) return (float)((double)v - q.pen_d);
        if ((double)v < -q.pen_d) return (float)((double)v + q.pen_d);
        return 0.f;
    }
    if (v > q.thresh) return (v - q.thresh) / q.denom;
    if (v < -q.thresh) return (v + q.thresh) / q.denom;
    return 0.f;
}
__device__ __forceinline__ float prox_active(float v, const WideProx& q)
{
    if (!q.enet) {
        if (v > q.pen_f) return v - q.pen_f;
        if (v < -q.pen_f) return v + q.pen_f;
        return 0.f;
    }
    if (v > q.thresh) return (v - q.thresh) / q.denom;
    if (v < -q.thresh) return (v + q.thresh) / q.denom;
    return 0.f;
}

// tmp = (Ax + z) + y / frho      [optionally / gamma]
__global__ void __launch_bounds__(WT) wide_tmp_kernel(const float* __restrict__ Ax, const float* __restrict__ z, const float* __restrict__ y,
                                                      float frho, float gamma, int divide, i64 n, float* __restrict__ tmp,
                                                      const WideCtl* __restrict__ ctl = nullptr)
{
    if (ctl) { if (ctl->stop) return; frho = ctl->frho; }
    const i64 i = (i64)blockIdx.x * WT + threadIdx.x;
    if (i >= n) return;
    float t = (Ax[i] + z[i]) + y[i] / frho;
    if (divide) t = t / gamma;
    tmp[i] = t;
}

// regular step, second half: x_j = prox(-vec_j / gamma + x_j) for every j
__global__ void __launch_bounds__(WT) wide_prox_all_kernel(const float* __restrict__ vec, float* __restrict__ x, i64 p, float gamma, WideProx q)
{
    const i64 j = (i64)blockIdx.x * WT + threadIdx.x;
    if (j >= p) return;
    const float v = -vec[j] / gamma + x[j];
    x[j] = prox_regular(v, q);
}

// active step: one warp per support column
__global__ void __launch_bounds__(WT) wide_active_kernel(const float* __restrict__ X, i64 ldx, i64 n, const float* __restrict__ tmp,
                                                         const int* __restrict__ supp, int nnz, float* __restrict__ x, WideProx q,
                                                         const WideCtl* __restrict__ ctl = nullptr)
{
    if (ctl) {                                           // batched run: support size and prox parameters live on the device
        if (ctl->stop) return;
        nnz = ctl->nnz; q.pen_f = ctl->pen_f; q.thresh = ctl->thresh; q.denom = ctl->denom;
    }
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * WT) >> 5;
    for (int k = (blockIdx.x * WT + threadIdx.x) >> 5; k < nnz; k += warps) {
        const int j = supp[k];
        const float* col = X + (i64)j * ldx;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        i64 i = lane;
        for (; i + 96 < n; i += 128) {
            s0 = fmaf(tmp[i], col[i], s0); s1 = fmaf(tmp[i + 32], col[i + 32], s1);
            s2 = fmaf(tmp[i + 64], col[i + 64], s2); s3 = fmaf(tmp[i + 96], col[i + 96], s3);
        }
        for (; i < n; i += 32) s0 = fmaf(tmp[i], col[i], s0);
        const float d = warp_sum((s0 + s1) + (s2 + s3));
        if (lane == 0) x[j] = prox_active(x[j] - d, q);
    }
}

// ---- screened regular step -----------------------------------------------------------------------------
// A regular step evaluates x_j = prox(-X_j' tmp / gamma + x_j) for ALL p columns although all but a few thousand of them
// come out zero: 4 n p bytes (40 GB at C3) per step, 413 times per path.  The screen reads an fp16 copy of the design
// (2 n p bytes) and proves most of them zero without touching the float32 data:
//     S_j = sum_i fp16(x_ij) tmp_i  (float32 accumulation),     |V_j - S_j| <= E_j
// where V_j is the value gemv_t computes and
//     E_j = [ (2^-11 + 2 g (1 + 2^-11)) |x_j|_2 + 2^-25 sqrt(n) ] |tmp|_2 * 1.001,    g = (n / 32 + 16) 2^-24
// (fp16 rounding of a normal number <= 2^-11 relative, of a subnormal <= 2^-25 absolute; g bounds the accumulated rounding
// of either float32 dot product: no sum runs over more than n / 32 + 16 sequential additions; Cauchy-Schwarz; 1.001 covers
// the float32 norms).  A column with x_j = 0 and (|S_j| + E_j) <= thr (1 - 1e-6), thr = gamma * (prox threshold), stays
// exactly zero in the unscreened computation too (|fl(-V_j / gamma)| <= threshold); every other column -- the current
// support and the few per mille near the threshold -- is recomputed from the float32 data with gemv_t's own summation
// order (gemv_t_list).  The step is therefore bit-identical to the unscreened one (tests: B200ADMM_WIDE_SCREEN=0).
__global__ void __launch_bounds__(WT) wide_half_copy_kernel(const float* __restrict__ X, i64 ldx, i64 n, i64 p, i64 ldh,
                                                            __half* __restrict__ Xh, float* __restrict__ colnorm)
{
    const int lane = threadIdx.x & 31;
    const i64 nwarps = (i64)gridDim.x * (WT >> 5);
    for (i64 j = (i64)blockIdx.x * (WT >> 5) + (threadIdx.x >> 5); j < p; j += nwarps) {
        const float* col = X + j * ldx;
        __half* dst = Xh + j * ldh;
        float s = 0.f;
        for (i64 i = lane; i < ldh; i += 32) {
            const float v = i < n ? col[i] : 0.f;
            dst[i] = __float2half_rn(v);
            s = fmaf(v, v, s);
        }
        s = warp_sum(s);
        if (lane == 0) colnorm[j] = sqrtf(s);
    }
}
// out[0] = |v|_2 (single CTA, fixed order)
__global__ void __launch_bounds__(1024) wide_norm2_kernel(const float* __restrict__ v, i64 n, float* __restrict__ out)
{
    __shared__ float scratch[33];
    float s = 0.f;
    for (i64 i = threadIdx.x; i < n; i += 1024) s = fmaf(v[i], v[i], s);
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) out[0] = sqrtf(s);
}
// mark[j] = 1 if column j must be evaluated exactly, 0 if it provably stays zero
__global__ void __launch_bounds__(WT) wide_screen_kernel(const __half* __restrict__ Xh, i64 ldh, i64 n, i64 p, const float* __restrict__ tmp,
                                                         const float* __restrict__ x, const float* __restrict__ colnorm,
                                                         const float* __restrict__ tmpnorm, float kx, float kn, float thr_safe,
                                                         float* __restrict__ mark)
{
    const int lane = threadIdx.x & 31;
    const i64 nwarps = (i64)gridDim.x * (WT >> 5);
    const float tn = *tmpnorm;
    const i64 nv = ldh / 8;                                  // 16-byte groups of 8 halves per column
    (void)n;
    for (i64 j = (i64)blockIdx.x * (WT >> 5) + (threadIdx.x >> 5); j < p; j += nwarps) {
        if (x[j] != 0.f) { if (lane == 0) mark[j] = 1.f; continue; }
        const uint4* col = reinterpret_cast<const uint4*>(Xh + j * ldh);
        float s0 = 0.f, s1 = 0.f;
        // eight halves per 16-byte load against two float4 of tmp (tmp is padded with zeros up to ldh by the caller);
        // four loads in flight per lane
        auto acc8 = [&](const uint4& h, i64 g) {
            const float4 t0 = __ldg(reinterpret_cast<const float4*>(tmp) + 2 * g), t1 = __ldg(reinterpret_cast<const float4*>(tmp) + 2 * g + 1);
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
            const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&h.z)), d = __half22float2(*reinterpret_cast<const __half2*>(&h.w));
            s0 = fmaf(a.x, t0.x, s0); s1 = fmaf(a.y, t0.y, s1); s0 = fmaf(b.x, t0.z, s0); s1 = fmaf(b.y, t0.w, s1);
            s0 = fmaf(c.x, t1.x, s0); s1 = fmaf(c.y, t1.y, s1); s0 = fmaf(d.x, t1.z, s0); s1 = fmaf(d.y, t1.w, s1);
        };
        auto ldh4 = [&](i64 g) {
            uint4 h;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(h.x), "=r"(h.y), "=r"(h.z), "=r"(h.w) : "l"(col + g));
            return h;
        };
        i64 g = lane;
        for (; g + 96 < nv; g += 128) {
            const uint4 h0 = ldh4(g), h1 = ldh4(g + 32), h2 = ldh4(g + 64), h3 = ldh4(g + 96);
            acc8(h0, g); acc8(h1, g + 32); acc8(h2, g + 64); acc8(h3, g + 96);
        }
        for (; g < nv; g += 32) acc8(ldh4(g), g);
        const float sj = warp_sum(s0 + s1);
        if (lane == 0) {
            const float e = (kx * colnorm[j] + kn) * tn;
            mark[j] = (fabsf(sj) + e <= thr_safe) ? 0.f : 1.f;
        }
    }
}
// x_j = prox(-vec_j / gamma + x_j) for the listed columns
__global__ void __launch_bounds__(WT) wide_prox_list_kernel(const float* __restrict__ vec, const int* __restrict__ list, int count,
                                                            float* __restrict__ x, float gamma, WideProx q)
{
    const int k = blockIdx.x * WT + threadIdx.x;
    if (k >= count) return;
    const int j = list[k];
    const float v = -vec[j] / gamma + x[j];
    x[j] = prox_regular(v, q);
}

// ---- fused active-set step: x-update and the new A x from ONE read of the support columns ------------------
// wide_active_kernel + wide_ax_kernel read every support column twice (dot product with tmp, then the row-parallel sum
// A x), both as gathers at about half the HBM rate.  Here a CTA owns a contiguous range of the support list; its 512
// threads hold a column in registers (thread t: rows 4 (t + 512 v) .. + 3), reduce the dot product across the CTA,
// apply the prox and add x_j times the SAME registers to the CTA's partial A x; the next column is loaded while the
// reduction of the current one runs.  part[b][0 .. n) = partial of logical CTA b; wide_zstep_kernel adds the
// min(SMs, nnz) partials in CTA order.  The partition depends on the support size only (not on the launch bound), so
// batched and host-driven runs stay bit-identical.
constexpr int FT = 512;
__host__ __device__ __forceinline__ int wide_fused_ctas(int nnz, int sms) { return nnz < sms ? nnz : sms; }
template <int NV>
__global__ void __launch_bounds__(FT, 1) wide_active_ax_kernel(const float* __restrict__ X, i64 ldx, i64 n, const float* __restrict__ tmp,
                                                               const int* __restrict__ supp, int nnz, int sms, float* __restrict__ x, WideProx q,
                                                               float* __restrict__ part, const WideCtl* __restrict__ ctl)
{
    __shared__ float scratch[33];
    if (ctl) {
        if (ctl->stop) return;
        nnz = ctl->nnz; q.pen_f = ctl->pen_f; q.thresh = ctl->thresh; q.denom = ctl->denom;
    }
    const int G = wide_fused_ctas(nnz, sms);
    if ((int)blockIdx.x >= G) return;
    const int per = (nnz + G - 1) / G;
    const int k0 = blockIdx.x * per, k1 = min(nnz, k0 + per);
    const int n4 = (int)(n / 4);                              // (n % 4 == 0: checked by the caller)
    const int tid = threadIdx.x;
    float4 t[NV], acc[NV], c[NV], cn[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const int idx = tid + v * FT;
        t[v] = idx < n4 ? __ldg(reinterpret_cast<const float4*>(tmp) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    auto load_col = [&](int k, float4 (&dst)[NV]) {
        const float4* col = reinterpret_cast<const float4*>(X + (i64)supp[k] * ldx);
#pragma unroll
        for (int v = 0; v < NV; v++) {
            const int idx = tid + v * FT;
            dst[v] = idx < n4 ? ld_stream_f4(col + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    if (k0 < k1) load_col(k0, c);
    for (int k = k0; k < k1; k++) {
        if (k + 1 < k1) load_col(k + 1, cn);
        float d = 0.f;
#pragma unroll
        for (int v = 0; v < NV; v++) {
            d = fmaf(c[v].x, t[v].x, d); d = fmaf(c[v].y, t[v].y, d); d = fmaf(c[v].z, t[v].z, d); d = fmaf(c[v].w, t[v].w, d);
        }
        d = block_sum(d, scratch);
        const int j = supp[k];
        const float xn = prox_active(x[j] - d, q);
        __syncthreads();                                      // everybody has read x[j]
        if (tid == 0) x[j] = xn;
        if (xn != 0.f) {
#pragma unroll
            for (int v = 0; v < NV; v++) {
                acc[v].x += c[v].x * xn; acc[v].y += c[v].y * xn; acc[v].z += c[v].z * xn; acc[v].w += c[v].w * xn;
            }
        }
#pragma unroll
        for (int v = 0; v < NV; v++) c[v] = cn[v];
    }
    float4* dst = reinterpret_cast<float4*>(part + (size_t)blockIdx.x * (size_t)n);
#pragma unroll
    for (int v = 0; v < NV; v++) {
        const int idx = tid + v * FT;
        if (idx < n4) dst[idx] = acc[v];
    }
}

// ---- device-side loop control for runs of active-set steps ---------------------------------------------------
// Between two regular steps (iteration counters 4^k - 1) every iteration is an active-set step whose launch shapes are
// bounded by the support size at the start of the run (the support only shrinks).  Such a run is enqueued as ONE batch
// without host round trips: the scalars the host loop computes after every iteration -- tolerances from the previous
// iterate, residuals, the stopping rule, ADMMBase::update_rho (src/ADMMBase.h:192-216) and the prox parameters derived
// from rho -- are evaluated by wide_control_kernel in the same double / float expressions (IEEE, no contraction on either
// side), and every kernel of the batch starts by looking at the stop flag.  The host reads the block back once per batch.
__device__ __forceinline__ void wide_derive_params(WideCtl* c, const WideCtlConst& k)
{
    c->frho = (float)c->rho;
    const double pen_d = k.lambda / (c->rho * (double)k.gamma);
    c->pen_f = (float)pen_d;
    c->thresh = k.alpha_f * c->pen_f;                                    // active_set_update(): all Scalar
    c->denom = (float)(1.0 + (double)c->pen_f * (1.0 - (double)k.alpha_f));
    c->den = (float)(-1 - c->rho);
}
__global__ void wide_ctl_init_kernel(WideCtl* c, WideCtlConst k, const int* __restrict__ nnz_dev)
{
    if (threadIdx.x == 0) {
        c->stop = 0; c->iters = 0; c->work_bytes = 0.0; c->nnz = *nnz_dev;
        wide_derive_params(c, k);
    }
}
// end of an iteration: sums -> tolerances, residuals, stopping rule, rho balancing, next parameters, trace row
__global__ void __launch_bounds__(1024) wide_control_kernel(WideCtl* c, WideCtlConst k, const float* __restrict__ psums, int nblocks,
                                                            const int* __restrict__ nnz_dev, double* __restrict__ trace, int trace_cap,
                                                            const float* __restrict__ Ax, const float* __restrict__ z, const float* __restrict__ y,
                                                            float* __restrict__ tmp)
{
    __shared__ double h[5];
    if (c->stop) return;
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q < 5) {
        double s = 0.0;
        for (int b = lane; b < nblocks; b += 32) s += (double)psums[(size_t)b * 5 + q];
        s = warp_sum(s);
        if (lane == 0) h[q] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
    const int i = k.i_start + c->iters;
    const double eps_primal = fmax((double)sqrtf((float)c->sAx2), (double)sqrtf((float)c->sz2)) * k.eps_rel + k.sqrt_n * k.eps_abs;
    const double eps_dual = (double)(k.sqrt_sprad * sqrtf((float)c->sy2)) * k.eps_rel + k.sqrt_p * k.eps_abs;
    const double rho = c->rho;
    const double resid_dual = rho * (double)k.sqrt_sprad * (double)sqrtf((float)h[0]);
    const double resid_primal = (double)sqrtf((float)h[1]);
    const int nnz_before = c->nnz, nnz = *nnz_dev;
    c->sAx2 = h[2]; c->sz2 = h[3]; c->sy2 = h[4];
    c->nnz = nnz;
    c->work_bytes += 4.0 * (double)k.n * (double)nnz_before + 4.0 * (double)k.n * (double)nnz + 48.0 * (double)k.n;
    c->eps_primal = eps_primal; c->resid_primal = resid_primal; c->eps_dual = eps_dual; c->resid_dual = resid_dual;
    if (trace && i < trace_cap) {
        double* row = trace + 5 * (size_t)i;
        row[0] = eps_primal; row[1] = resid_primal; row[2] = eps_dual; row[3] = resid_dual; row[4] = rho;
    }
    c->iters += 1;
    if (resid_primal < eps_primal && resid_dual < eps_dual) c->stop = 1;
    else {
        if (i > 3) {
            double r = rho;                                                   // balance_rho
            if (resid_primal / eps_primal > 10 * resid_dual / eps_dual) r *= 2;
            else if (resid_dual / eps_dual > 10 * resid_primal / eps_primal) r /= 2;
            if (resid_primal < eps_primal) r /= 1.2;
            if (resid_dual < eps_dual) r *= 1.2;
            c->rho = r;
        }
        wide_derive_params(c, k);
    }
    }
    __syncthreads();
    // tmp of the NEXT active-set step with the new rho (what wide_tmp_kernel computes: one launch less per iteration)
    if (c->stop || tmp == nullptr) return;
    const float frho = c->frho;
    for (long long i = threadIdx.x; i < k.n; i += blockDim.x) tmp[i] = ((Ax[i] + z[i]) + y[i] / frho) / k.gamma;
}

// single-CTA stable compaction for short candidate lists (<= a few thousand entries: the active-set steps): the three
// launches of the general scheme cost more in launch latency than in work
__global__ void __launch_bounds__(1024) compact_small_kernel(const int* __restrict__ src, const float* __restrict__ x, int len,
                                                             int* __restrict__ out, int* __restrict__ total, const WideCtl* __restrict__ ctl)
{
    if (ctl) { if (ctl->stop) return; len = ctl->nnz; }
    __shared__ int wsum[32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < len; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        int keep = 0, j = 0;
        if (i < len) { j = src[i]; keep = x[j] != 0.f; }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(ballot);
        __syncthreads();
        int chunk_total = 0;
        if (warp == 0) {
            int v = wsum[lane];
            const int mine = v;
            for (int off = 1; off < 32; off <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, off); if (lane >= off) v += t; }
            wsum[lane] = v - mine;                                   // exclusive warp offsets
            chunk_total = __shfl_sync(0xffffffffu, v, 31);
        }
        __syncthreads();
        if (keep) out[base + wsum[warp] + __popc(ballot & ((1u << lane) - 1u))] = j;
        __syncthreads();
        if (threadIdx.x == 0) base += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = base;
}

// ---- stable compaction of "x[j] != 0" into a sorted index list ---------------------------------
// src == nullptr: candidates are 0..len-1; otherwise candidates are src[0..len-1] (already sorted)
__global__ void __launch_bounds__(1024) compact_count_kernel(const int* __restrict__ src, const float* __restrict__ x, int len, int* __restrict__ counts,
                                                             const WideCtl* __restrict__ ctl = nullptr)
{
    if (ctl) { if (ctl->stop) return; len = ctl->nnz; }   // (blocks beyond the exact length count zero)
    const int i = blockIdx.x * 1024 + threadIdx.x;
    int keep = 0;
    if (i < len) { const int j = src ? src[i] : i; keep = x[j] != 0.f; }
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) counts[blockIdx.x] = c;
}
__global__ void __launch_bounds__(1024) compact_scan_kernel(int* __restrict__ counts, int nb, int* __restrict__ total,
                                                            const WideCtl* __restrict__ ctl = nullptr)
{
    if (ctl && ctl->stop) return;
    __shared__ int sh[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? counts[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            int t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb) counts[i] = carry + sh[threadIdx.x] - v;      // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(1024) compact_scatter_kernel(const int* __restrict__ src, const float* __restrict__ x, int len,
                                                               const int* __restrict__ offsets, int* __restrict__ out,
                                                               const WideCtl* __restrict__ ctl = nullptr)
{
    if (ctl) { if (ctl->stop) return; len = ctl->nnz; }
    __shared__ int wsum[32];
    const int i = blockIdx.x * 1024 + threadIdx.x;
    int keep = 0, j = 0;
    if (i < len) { j = src ? src[i] : i; keep = x[j] != 0.f; }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wsum[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        int v = wsum[lane];
        for (int off = 1; off < 32; off <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, off); if (lane >= off) v += t; }
        wsum[lane] = v - wsum[lane];                                 // exclusive warp offsets
    }
    __syncthreads();
    if (keep) out[offsets[blockIdx.x] + wsum[warp] + __popc(ballot & ((1u << lane) - 1u))] = j;
}

// ---- z step ----------------------------------------------------------------------------------------
// partial Ax over a chunk of the support: part[chunk][i] = sum_k X(i, supp_k) * x[supp_k]
// chunking of the support for the partial Ax sums (the same on host and device)
__host__ __device__ __forceinline__ int wide_chunks(int nnz, int max_chunks)
{
    return nnz > 0 ? max(1, min(max_chunks, (nnz + 255) / 256)) : 0;
}
__global__ void __launch_bounds__(WT) wide_ax_kernel(const float* __restrict__ X, i64 ldx, i64 n, const int* __restrict__ supp,
                                                     const int* __restrict__ nnz_dev, int max_chunks,
                                                     const float* __restrict__ x, float* __restrict__ part,
                                                     const WideCtl* __restrict__ ctl = nullptr)
{
    if (ctl && ctl->stop) return;
    __shared__ int sj[WT];
    __shared__ float sv[WT];
    const int nnz = *nnz_dev;                             // exact; the grid was sized from an upper bound
    const int chunks = wide_chunks(nnz, max_chunks);
    if ((int)blockIdx.y >= chunks) return;
    const int per_chunk = (nnz + chunks - 1) / chunks;
    const i64 i = (i64)blockIdx.x * WT + threadIdx.x;
    const int k0 = blockIdx.y * per_chunk, k1 = min(nnz, k0 + per_chunk);
    float acc = 0.f;
    for (int kb = k0; kb < k1; kb += WT) {
        const int cnt = min(WT, k1 - kb);
        __syncthreads();
        if ((int)threadIdx.x < cnt) { const int j = supp[kb + threadIdx.x]; sj[threadIdx.x] = j; sv[threadIdx.x] = x[j]; }
        __syncthreads();
        if (i < n) {
            int k = 0;
            for (; k + 4 <= cnt; k += 4) {
                const float a0 = X[i + (i64)sj[k] * ldx], a1 = X[i + (i64)sj[k + 1] * ldx];
                const float a2 = X[i + (i64)sj[k + 2] * ldx], a3 = X[i + (i64)sj[k + 3] * ldx];
                acc += a0 * sv[k]; acc += a1 * sv[k + 1]; acc += a2 * sv[k + 2]; acc += a3 * sv[k + 3];
            }
            for (; k < cnt; k++) acc += X[i + (i64)sj[k] * ldx] * sv[k];
        }
    }
    if (i < n) part[(i64)blockIdx.y * n + i] = acc;
}

// Ax = sum of chunks (fixed order); z = (ydat + y + frho Ax) / den; r = Ax + z; y += frho r; norms
__global__ void __launch_bounds__(WT) wide_zstep_kernel(const float* __restrict__ part, const int* __restrict__ nnz_dev, int max_chunks, i64 n,
                                                        const float* __restrict__ ydat,
                                                        float frho, float den, float* __restrict__ Ax, float* __restrict__ z,
                                                        float* __restrict__ y, float* __restrict__ psums,
                                                        const WideCtl* __restrict__ ctl = nullptr, int fused_nnz = -1, int sms = 0)
{
    if (ctl) { if (ctl->stop) return; frho = ctl->frho; den = ctl->den; if (fused_nnz >= 0) fused_nnz = ctl->nnz; }
    // after a fused active-set step the partials are those of its logical CTAs (support size BEFORE the step)
    const int chunks = fused_nnz >= 0 ? wide_fused_ctas(fused_nnz, sms) : wide_chunks(*nnz_dev, max_chunks);
    const i64 i = (i64)blockIdx.x * WT + threadIdx.x;
    float ps[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (i < n) {
        float ax = 0.f;
        for (int c = 0; c < chunks; c++) ax += part[(i64)c * n + i];
        const float yo = y[i];
        const float zn = ((ydat[i] + yo) + frho * ax) / den;
        const float dz = zn - z[i];
        const float r = ax + zn;
        const float yn = yo + frho * r;
        Ax[i] = ax; z[i] = zn; y[i] = yn;
        ps[0] = dz * dz; ps[1] = r * r; ps[2] = ax * ax; ps[3] = zn * zn; ps[4] = yn * yn;
    }
    __shared__ float s_red[WT / 32][5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 5; q++) ps[q] = warp_sum(ps[q]);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 5; q++) s_red[warp][q] = ps[q];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        float s = 0.f;
        for (int w = 0; w < WT / 32; w++) s += s_red[w][threadIdx.x];
        psums[(size_t)blockIdx.x * 5 + threadIdx.x] = s;
    }
}
__global__ void wide_finish_sums_kernel(const float* __restrict__ psums, int nblocks, const int* __restrict__ nnz, double* __restrict__ out6)
{
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q < 5) {
        double s = 0.0;
        for (int b = lane; b < nblocks; b += 32) s += (double)psums[(size_t)b * 5 + q];
        s = warp_sum(s);
        if (lane == 0) out6[q] = s;
    } else if (q == 5 && lane == 0) out6[5] = (double)*nnz;
}

// vals[i] = x[supp[i]]: the solution of one lambda as (index, value) pairs
__global__ void __launch_bounds__(WT) wide_gather_kernel(const float* __restrict__ x, const int* __restrict__ supp, int nnz, float* __restrict__ vals)
{
    const int i = blockIdx.x * WT + threadIdx.x;
    if (i < nnz) vals[i] = x[supp[i]];
}

// DataStd::recover + dgCMatrix assembly from per-lambda support lists (the sparse twin of finish_lasso_path:
// same arithmetic on the stored entries in index order, /root/reference/src/DataStd.h:183-207, Lasso.cpp:22-30)
void finish_sparse_path(std::vector<std::vector<int>>& idx, std::vector<std::vector<float>>& val, int nl, i64 p, int flag,
                        const std::vector<float>& meanX, const std::vector<float>& scaleX, float meanY, float scaleY, b200admm_path* out)
{
    std::vector<float> beta0(nl, 0.f);
    size_t nnz = 0;
    for (int k = 0; k < nl; k++) {
        float s = 0.f;
        for (size_t e = 0; e < idx[k].size(); e++) {
            const int j = idx[k][e];
            float c = val[k][e];
            if (flag == 1 || flag == 3) c /= scaleX[j];
            if (flag != 0) c *= scaleY;
            if (flag == 2 || flag == 3) s += c * meanX[j];
            val[k][e] = c;
            if (c != 0.f) nnz++;
        }
        if (flag == 2 || flag == 3) beta0[k] = meanY - s;
        nnz++;                                              // the intercept row is always stored
    }
    out->colptr = (int64_t*)malloc(sizeof(int64_t) * (nl + 1));
    out->rowidx = (int*)malloc(sizeof(int) * std::max<size_t>(nnz, 1));
    out->val = (double*)malloc(sizeof(double) * std::max<size_t>(nnz, 1));
    if (!out->colptr || !out->rowidx || !out->val) throw CodeError(B200ADMM_ENOMEM, "out of host memory");
    size_t pos = 0;
    for (int k = 0; k < nl; k++) {
        out->colptr[k] = (int64_t)pos;
        out->rowidx[pos] = 0; out->val[pos] = (double)beta0[k]; pos++;
        for (size_t e = 0; e < idx[k].size(); e++)
            if (val[k][e] != 0.f) { out->rowidx[pos] = idx[k][e] + 1; out->val[pos] = (double)val[k][e]; pos++; }
    }
    out->colptr[nl] = (int64_t)pos;
    out->nrow = p + 1;
}

inline bool is_regular_update(unsigned c)
{
    if (c == 0 || c == 3 || c == 15 || c == 63) return true;
    c++;
    if (c & (c - 1)) return false;
    return (c & 0x55555555u) != 0;
}
inline void balance_rho(double& rho, double rp, double ep, double rd, double ed)
{
    if (rp / ep > 10 * rd / ed) rho *= 2;
    else if (rd / ed > 10 * rp / ep) rho /= 2;
    if (rp < ep) rho /= 1.2;
    if (rd < ed) rho *= 1.2;
}

// Y (cols x rows, column-major, leading dimension ldy) = X' for X (rows x cols, ldx): 32 x 32 tiles
// through shared memory, coalesced on both sides
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ X, i64 rows, i64 cols, i64 ldx, float* __restrict__ Y, i64 ldy)
{
    __shared__ float tile[32][33];
    const i64 r0 = (i64)blockIdx.x * 32, c0 = (i64)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int c = ty; c < 32; c += 8) {
        const i64 r = r0 + tx, cc = c0 + c;
        tile[c][tx] = (r < rows && cc < cols) ? X[r + cc * ldx] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const i64 cc = c0 + tx, rr = r0 + r;
        if (cc < cols && rr < rows) Y[cc + rr * ldy] = tile[tx][r];
    }
}

// C (n x n, full) = X X' accumulated over column chunks so that no float sum runs over more than
// 8192 terms before it is folded into C (keeps the float32 result unbiased for p ~ 1e6)
void gram_nt_chunked(cudaStream_t s, const float* X, i64 n, i64 p, i64 ldx, float* C)
{
    const i64 chunk = 8192;
    for (i64 k0 = 0; k0 < p; k0 += chunk) {
        const i64 kc = std::min(chunk, p - k0);
        const bool last = k0 + kc >= p;
        gemm<float>(s, false, true, n, n, kc, 1.f, X + k0 * ldx, ldx, X + k0 * ldx, ldx, k0 == 0 ? 0.f : 1.f, C, n,
                    GEMM_LOWER | (last ? GEMM_MIRROR : 0));
    }
}

}  // namespace

void solve_wide(const LassoRequest& rq, b200admm_path* out)
{
    const b200admm_data* d = rq.d;
    Context& c = ctx();
    cudaStream_t s = c.stream;
    const i64 n = d->n, p = d->p;
    if (d->dtype == B200ADMM_F64_DEVICE) throw ArgError("lasso / enet compute in float32: pass f64 host, f32 host or f32 device data");
    if (n < 3) throw CodeError(B200ADMM_ELANCZOS, "coarse eigenvalue estimate needs at least 3 observations (Spectra: 1 <= nev < ncv <= n)");
    const double t_begin = wall_now();
    const int flag = (rq.standardize ? 1 : 0) + (rq.intercept ? 2 : 0);
    EventTimer tm(s);
    b200admm_timing T;
    memset(&T, 0, sizeof T);

    // ---- ingest + DataStd (Lasso.cpp:45-68) -----------------------------------------------------
    DevBuf<float> Xs((size_t)n * (size_t)p), ydat(n);
    const float* X_in = Xs.p;
    tm.start();
    if (d->dtype == B200ADMM_F32_DEVICE) {
        X_in = (const float*)d->x;
        CUDA_CHECK(cudaMemcpyAsync(ydat.p, d->y, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    } else {
        ingest_f32(s, d->x, d->dtype, (size_t)n * (size_t)p, Xs.p);
        ingest_f32(s, d->y, d->dtype, (size_t)n, ydat.p);
    }
    T.ingest = tm.stop();
    DevBuf<float> d_meanX(p), d_scaleX(p), tmpv(2 * p + 8), y2(2);
    tm.start();
    d_meanX.zero(s);
    CUDA_CHECK(cudaMemsetAsync(y2.p, 0, 2 * sizeof(float), s));
    standardize_y<float>(s, ydat.p, n, flag, y2.p, tmpv.p);
    standardize_columns<float>(s, X_in, Xs.p, n, p, n, flag, d_meanX.p, d_scaleX.p, tmpv.p);
    std::vector<float> meanX(p, 0.f), scaleX(p, 1.f);
    float h2[2] = {0.f, 1.f};
    if (flag >= 2) CUDA_CHECK(cudaMemcpyAsync(meanX.data(), d_meanX.p, p * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (flag & 1) CUDA_CHECK(cudaMemcpyAsync(scaleX.data(), d_scaleX.p, p * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (flag) CUDA_CHECK(cudaMemcpyAsync(h2, y2.p, 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    T.standardize = tm.stop();
    const float meanY = flag >= 2 ? h2[0] : 0.f, scaleY = flag ? h2[1] : 1.f;
    const float* X = Xs.p;
    const i64 ldx = n;

    // ---- ctor: lambda0 = max |X'y|, gamma = coarse lambda_max(XX') (ADMMLassoWide.h:189-212) ------
    DevBuf<float> vec(p);
    tm.start();
    gemv_t<float>(s, X, n, p, ldx, ydat.p, vec.p);
    std::vector<float> h_xy(p);
    CUDA_CHECK(cudaMemcpyAsync(h_xy.data(), vec.p, p * sizeof(float), cudaMemcpyDeviceToHost, s));
    float sprad = 0.f;
    {
        DevBuf<float> Gn((size_t)n * (size_t)n);
        // XX' is the Gram matrix of Y = X' (p x n, its columns are the rows of X): transposing once turns
        // the product into the K-major shape the tcgen05 3xTF32 kernel takes (fp32-accurate, ~5x the
        // CUDA-core rate); shapes it declines fall back to the chunked CUDA-core product
        bool on_tensor = false;
        const char* gram_env = getenv("B200ADMM_GRAM");
        if (!(gram_env && !strcmp(gram_env, "simt")) && n >= 512) {      // small problems: plain fp32 product
            const i64 ldy = (p + 3) & ~(i64)3;
            DevBuf<float> Y((size_t)ldy * (size_t)n);
            if (ldy != p) Y.zero(s);
            dim3 tg((unsigned)((n + 31) / 32), (unsigned)((p + 31) / 32));
            if (tg.y <= 65535) {
                transpose_kernel<<<tg, 256, 0, s>>>(X, n, p, ldx, Y.p, ldy);
                KERNEL_CHECK();
                const int split = (flag == 3 && !gram_env) ? GRAM_SPLIT_F16 : GRAM_SPLIT_TF32;   // |x_std| <= sqrt(n) under flag 3
                on_tensor = gram_tn_tensor(s, Y.p, p, ldy, n, Gn.p, n, split);
                if (on_tensor && split == GRAM_SPLIT_F16 && gram_f16_overflowed(s))
                    throw CudaError("fp16 Gram split: a standardised value exceeds sqrt(n)");
                CUDA_CHECK(cudaStreamSynchronize(s));
            }
        }
        if (!on_tensor) gram_nt_chunked(s, X, n, p, ldx, Gn.p);
        T.gram = tm.stop();
        tm.start();
        sprad = coarse_eig_device(s, Gn.p, n, n, nullptr);
        T.eig = tm.stop();
    }
    float lambda0 = 0.f;
    for (i64 j = 0; j < p; j++) lambda0 = std::max(lambda0, std::fabs(h_xy[j]));
    const float alpha_f = (float)rq.alpha;
    if (rq.enet) lambda0 = (float)((double)lambda0 / ((double)alpha_f + 0.0001));

    // ---- lambda sequence ---------------------------------------------------------------------------
    std::vector<double> lam;
    if (rq.nlambda_given < 1) {
        if (rq.nlambda < 1) throw ArgError("nlambda must be at least 1");
        const double lmax = (double)lambda0 / (double)n * (double)scaleY;
        make_lambda_grid(lmax, rq.lmin_ratio > 0 ? rq.lmin_ratio : (n < p ? 0.01 : 1e-4), rq.nlambda, lam);
    } else {
        if (!rq.lambda_given) throw ArgError("lambda is null");
        lam.assign(rq.lambda_given, rq.lambda_given + rq.nlambda_given);
    }
    const int nl = (int)lam.size();

    // ---- fp16 copy + column norms for the screened regular steps (see wide_screen_kernel) -----------------------
    const char* screen_env = getenv("B200ADMM_WIDE_SCREEN");
    bool screen = !(screen_env && !strcmp(screen_env, "0")) && gemv_t_uses_warp_kernel(n) && p >= 4096 && p < 2147483647LL;
    const i64 ldh = (n + 7) & ~(i64)7;
    DevBuf<__half> Xh;
    DevBuf<float> colnorm, tmpnorm;
    DevBuf<int> cand;
    if (screen) {
        // (the half-precision copy is an optimisation: a design that leaves no room for it runs unscreened)
        try { Xh.alloc((size_t)ldh * (size_t)p); }
        catch (const CodeError&) { screen = false; }
    }
    if (screen) {
        tm.start();
        colnorm.alloc(p); tmpnorm.alloc(1); cand.alloc(p);
        wide_half_copy_kernel<<<(unsigned)std::min<i64>((p + 7) / 8, (i64)sm_count() * 16), WT, 0, s>>>(X, ldx, n, p, ldh, Xh.p, colnorm.p);
        KERNEL_CHECK();
        T.factor = tm.stop();                                   // (reported under `factor`: the wide path has no factorisation)
    }
    const double g_round = ((double)n / 32.0 + 16.0) * std::ldexp(1.0, -24);
    const float screen_kx = (float)((std::ldexp(1.0, -11) + 2.0 * g_round * (1.0 + std::ldexp(1.0, -11))) * 1.001);
    const float screen_kn = (float)(std::ldexp(1.0, -25) * std::sqrt((double)n) * 1.001);
    double screened_steps = 0, screened_cand = 0;
    int last_ncand = 0;

    // ---- state ---------------------------------------------------------------------------------------
    DevBuf<float> x(p), Ax(n), z(n), y(n), tmp(ldh);            // (tmp: zero-padded to the fp16 columns' length for the screen)
    tmp.zero(s);
    DevBuf<int> supp[2], nnz_dev(1), counts((size_t)((p + 1023) / 1024 + 1));
    supp[0].alloc(p); supp[1].alloc(p);
    x.zero(s); Ax.zero(s); z.zero(s); y.zero(s); nnz_dev.zero(s);
    const int zblocks = (int)((n + WT - 1) / WT);
    DevBuf<float> psums((size_t)zblocks * 5);
    DevBuf<double> sums6(6);
    int max_chunks = 64;
    const int sms_n = sm_count();
    DevBuf<float> part((size_t)std::max(max_chunks, sms_n) * (size_t)n);
    // fused active-set step (x-update + A x from one read of the support columns): columns of up to 12288 floats, 16-byte aligned
    const char* fuse_env = getenv("B200ADMM_WIDE_FUSE");
    const bool fused_ok = !(fuse_env && !strcmp(fuse_env, "0")) && n % 4 == 0 && n <= 4 * FT * 6 && (((uintptr_t)X) & 15) == 0;
    auto launch_fused = [&](int grid, const int* supp_p, int nnz_arg, const WideProx& qq, const WideCtl* cp) {
        const int nv = (int)((n / 4 + FT - 1) / FT);
        if (nv <= 2) wide_active_ax_kernel<2><<<grid, FT, 0, s>>>(X, ldx, n, tmp.p, supp_p, nnz_arg, sms_n, x.p, qq, part.p, cp);
        else if (nv <= 4) wide_active_ax_kernel<4><<<grid, FT, 0, s>>>(X, ldx, n, tmp.p, supp_p, nnz_arg, sms_n, x.p, qq, part.p, cp);
        else wide_active_ax_kernel<6><<<grid, FT, 0, s>>>(X, ldx, n, tmp.p, supp_p, nnz_arg, sms_n, x.p, qq, part.p, cp);
        KERNEL_CHECK();
    };
    int cur_supp = 0, nnz = 0, nnz_bound = 0;
    // runs of active-set steps are enqueued as batches with device-side loop control (B200ADMM_WIDE_BATCH=0: host-driven)
    const char* batch_env = getenv("B200ADMM_WIDE_BATCH");
    const bool batched = !(batch_env && !strcmp(batch_env, "0"));
    DevBuf<WideCtl> ctl(1);
    DevBuf<double> trace_dev;

    auto compact = [&](const int* src, int len, int* dst) {
        const int nb = (len + 1023) / 1024;
        if (nb == 0) { CUDA_CHECK(cudaMemsetAsync(nnz_dev.p, 0, sizeof(int), s)); return; }
        compact_count_kernel<<<nb, 1024, 0, s>>>(src, x.p, len, counts.p); KERNEL_CHECK();
        compact_scan_kernel<<<1, 1024, 0, s>>>(counts.p, nb, nnz_dev.p); KERNEL_CHECK();
        compact_scatter_kernel<<<nb, 1024, 0, s>>>(src, x.p, len, counts.p, dst); KERNEL_CHECK();
    };

    TraceRequest& tr = trace_request();
    std::vector<std::vector<int>> idx_all(nl);
    std::vector<std::vector<float>> val_all(nl);
    out->niter = (int*)malloc(sizeof(int) * nl);
    out->lambda = (double*)malloc(sizeof(double) * nl);
    if (!out->niter || !out->lambda) throw CodeError(B200ADMM_ENOMEM, "out of host memory");

    double rho = rq.opts.rho;
    double sAx2 = 0, sz2 = 0, sy2 = 0;                      // squared norms of the current Ax, z, y
    const double eps_abs = rq.opts.eps_abs, eps_rel = rq.opts.eps_rel;
    const float gamma = sprad;
    const float sqrt_sprad = std::sqrt(sprad);
    // algorithmic bytes of the iteration phase: a regular step reads all of X (4 n p), an
    // active-set step the nnz_k support columns (4 n nnz_k); the z step reads the nnz_{k+1} columns of the new
    // support; 12 n-vectors of traffic besides.  Data dependent, so it is accumulated while the loop runs.
    double work_bytes = 0, work_regular = 0, work_active = 0;
    tm.start();
    for (int k = 0; k < nl; k++) {
        const float lambda = (float)(lam[k] * (double)n / (double)scaleY);      // Lasso.cpp:99, stored as Scalar
        if (k == 0) {
            if (rho <= 0) rho = std::pow((double)(lambda / sprad), 1.0 / 3);    // ADMMLassoWide.h:228-229
        }
        unsigned iter_counter = 0;                                                // init / init_warm
        const bool tracing = tr.buf && tr.cap > 0 && tr.which == k;
        if (tracing && batched && !trace_dev.p) trace_dev.alloc((size_t)5 * tr.cap);
        int i;
        for (i = 0; i < rq.opts.maxit; ) {
            // ---------------- a run of active-set steps as one batch (device-side loop control) ----------------
            {
                const bool shortcut = !rq.enet && (double)lambda > (double)lambda0 - 1e-5;
                auto regular_at = [&](unsigned c) { return is_regular_update(c) && (!rq.enet || lambda < lambda0); };
                if (batched && !shortcut && nnz > 0 && !regular_at(iter_counter)) {
                    int cnt = 0;
                    while (cnt < 64 && i + cnt < rq.opts.maxit && !regular_at(iter_counter + (unsigned)cnt)) cnt++;
                    WideCtl hc;
                    memset(&hc, 0, sizeof hc);
                    hc.rho = rho; hc.sAx2 = sAx2; hc.sz2 = sz2; hc.sy2 = sy2;
                    CUDA_CHECK(cudaMemcpyAsync(ctl.p, &hc, sizeof hc, cudaMemcpyHostToDevice, s));
                    WideCtlConst kc;
                    kc.eps_abs = eps_abs; kc.eps_rel = eps_rel; kc.sqrt_n = std::sqrt((double)n); kc.sqrt_p = std::sqrt((double)p);
                    kc.lambda = (double)lambda; kc.gamma = gamma; kc.sqrt_sprad = sqrt_sprad; kc.alpha_f = alpha_f;
                    kc.enet = rq.enet ? 1 : 0; kc.i_start = i; kc.n = (long long)n;
                    wide_ctl_init_kernel<<<1, 32, 0, s>>>(ctl.p, kc, nnz_dev.p); KERNEL_CHECK();
                    WideProx q;
                    q.enet = rq.enet ? 1 : 0; q.pen_d = 0; q.pen_f = 0; q.thresh = 0; q.denom = 1;     // (parameters come from the control block)
                    const int ablocks = std::min((nnz + 7) / 8, sm_count() * 8);
                    const int nb = (nnz + 1023) / 1024;
                    const int chunks_bound = wide_chunks(nnz, max_chunks);
                    int cs = cur_supp;
                    // five launches per iteration: x-update, compaction, Ax partials, z-step, control (+ the next tmp)
                    wide_tmp_kernel<<<zblocks, WT, 0, s>>>(Ax.p, z.p, y.p, 0.f, gamma, 1, n, tmp.p, ctl.p); KERNEL_CHECK();
                    for (int b = 0; b < cnt; b++) {
                        if (fused_ok) launch_fused(std::min(sms_n, nnz), supp[cs].p, nnz, q, ctl.p);
                        else { wide_active_kernel<<<ablocks, WT, 0, s>>>(X, ldx, n, tmp.p, supp[cs].p, nnz, x.p, q, ctl.p); KERNEL_CHECK(); }
                        if (nnz <= 8192) {
                            compact_small_kernel<<<1, 1024, 0, s>>>(supp[cs].p, x.p, nnz, supp[cs ^ 1].p, nnz_dev.p, ctl.p); KERNEL_CHECK();
                        } else {
                            compact_count_kernel<<<nb, 1024, 0, s>>>(supp[cs].p, x.p, nnz, counts.p, ctl.p); KERNEL_CHECK();
                            compact_scan_kernel<<<1, 1024, 0, s>>>(counts.p, nb, nnz_dev.p, ctl.p); KERNEL_CHECK();
                            compact_scatter_kernel<<<nb, 1024, 0, s>>>(supp[cs].p, x.p, nnz, counts.p, supp[cs ^ 1].p, ctl.p); KERNEL_CHECK();
                        }
                        cs ^= 1;
                        if (fused_ok) {
                            wide_zstep_kernel<<<zblocks, WT, 0, s>>>(part.p, nnz_dev.p, max_chunks, n, ydat.p, 0.f, 0.f, Ax.p, z.p, y.p, psums.p, ctl.p, 0, sms_n);
                            KERNEL_CHECK();
                        } else {
                            wide_ax_kernel<<<dim3((unsigned)zblocks, (unsigned)chunks_bound), WT, 0, s>>>(X, ldx, n, supp[cs].p, nnz_dev.p, max_chunks, x.p, part.p, ctl.p);
                            KERNEL_CHECK();
                            wide_zstep_kernel<<<zblocks, WT, 0, s>>>(part.p, nnz_dev.p, max_chunks, n, ydat.p, 0.f, 0.f, Ax.p, z.p, y.p, psums.p, ctl.p); KERNEL_CHECK();
                        }
                        wide_control_kernel<<<1, 1024, 0, s>>>(ctl.p, kc, psums.p, zblocks, nnz_dev.p, tracing ? trace_dev.p : nullptr, tracing ? tr.cap : 0,
                                                               Ax.p, z.p, y.p, b + 1 < cnt ? tmp.p : nullptr);
                        KERNEL_CHECK();
                    }
                    CUDA_CHECK(cudaMemcpyAsync(&hc, ctl.p, sizeof hc, cudaMemcpyDeviceToHost, s));
                    CUDA_CHECK(cudaStreamSynchronize(s));
                    const int done = hc.iters;
                    if (tracing && done > 0 && i < tr.cap) {
                        const int rows = std::min(done, tr.cap - i);
                        CUDA_CHECK(cudaMemcpy(tr.buf + 5 * (size_t)i, trace_dev.p + 5 * (size_t)i, sizeof(double) * 5 * rows, cudaMemcpyDeviceToHost));
                        if (tr.nrows) *tr.nrows = i + rows;
                    }
                    rho = hc.rho; sAx2 = hc.sAx2; sz2 = hc.sz2; sy2 = hc.sy2; nnz = hc.nnz; nnz_bound = nnz;
                    cur_supp ^= (done & 1);
                    iter_counter += (unsigned)done;
                    work_bytes += hc.work_bytes; work_active += done;
                    if (hc.stop) { i += done - 1; break; }               // converged at iteration index i + done - 1
                    i += done;
                    continue;
                }
            }
            const double eps_primal = std::max((double)std::sqrt((float)sAx2), (double)std::sqrt((float)sz2)) * eps_rel + std::sqrt((double)n) * eps_abs;
            const double eps_dual = (double)(sqrt_sprad * std::sqrt((float)sy2)) * eps_rel + std::sqrt((double)p) * eps_abs;
            const float frho = (float)rho;
            const int nnz_before = nnz;
            int step_kind = 0;                                             // 0: x = 0 shortcut, 1: regular, 2: active set
            bool fused_step = false;                                       // this iteration's A x partials come from the fused active-set kernel
            // ---------------- x step ----------------
            if (!rq.enet && (double)lambda > (double)lambda0 - 1e-5) {
                if (nnz > 0) { x.zero(s); nnz = 0; CUDA_CHECK(cudaMemsetAsync(nnz_dev.p, 0, sizeof(int), s)); }
                nnz_bound = 0;
            } else {
                WideProx q;
                q.enet = rq.enet ? 1 : 0;
                q.pen_d = (double)lambda / (rho * (double)gamma);
                q.pen_f = (float)q.pen_d;
                const bool regular = is_regular_update(iter_counter) && (!rq.enet || lambda < lambda0);
                if (regular) {
                    q.thresh = (float)((double)alpha_f * q.pen_d);                 // enet(): Scalar thresh = alpha * penalty(double)
                    q.denom = (float)(1.0 + q.pen_d * (1.0 - (double)alpha_f));
                    wide_tmp_kernel<<<zblocks, WT, 0, s>>>(Ax.p, z.p, y.p, frho, gamma, 0, n, tmp.p); KERNEL_CHECK();
                    if (screen) {
                        // threshold on |X_j' tmp| below which prox(-vec_j / gamma) = 0: gamma * pen (lasso) / gamma * thresh (enet)
                        const double thr = (double)gamma * (q.enet ? (double)q.thresh : q.pen_d);
                        const float thr_safe = (float)(thr * (1.0 - 1e-6));
                        wide_norm2_kernel<<<1, 1024, 0, s>>>(tmp.p, n, tmpnorm.p); KERNEL_CHECK();
                        wide_screen_kernel<<<(unsigned)std::min<i64>((p + 7) / 8, (i64)sm_count() * 16), WT, 0, s>>>(
                            Xh.p, ldh, n, p, tmp.p, x.p, colnorm.p, tmpnorm.p, screen_kx, screen_kn, thr_safe, vec.p);
                        KERNEL_CHECK();
                        // candidates = marked columns (sorted); their count comes to the host (one small read per regular step)
                        {
                            const int nb = (int)((p + 1023) / 1024);
                            compact_count_kernel<<<nb, 1024, 0, s>>>(nullptr, vec.p, (int)p, counts.p); KERNEL_CHECK();
                            compact_scan_kernel<<<1, 1024, 0, s>>>(counts.p, nb, nnz_dev.p); KERNEL_CHECK();
                            compact_scatter_kernel<<<nb, 1024, 0, s>>>(nullptr, vec.p, (int)p, counts.p, cand.p); KERNEL_CHECK();
                        }
                        int ncand = 0;
                        CUDA_CHECK(cudaMemcpyAsync(&ncand, nnz_dev.p, sizeof(int), cudaMemcpyDeviceToHost, s));
                        CUDA_CHECK(cudaStreamSynchronize(s));
                        screened_steps += 1; screened_cand += ncand; last_ncand = ncand;
                        gemv_t_list<float>(s, X, n, ldx, tmp.p, cand.p, ncand, vec.p);
                        if (ncand > 0) {
                            wide_prox_list_kernel<<<(unsigned)((ncand + WT - 1) / WT), WT, 0, s>>>(vec.p, cand.p, ncand, x.p, gamma, q); KERNEL_CHECK();
                        }
                        compact(cand.p, ncand, supp[cur_supp].p);
                    } else {
                        gemv_t<float>(s, X, n, p, ldx, tmp.p, vec.p);
                        wide_prox_all_kernel<<<(unsigned)((p + WT - 1) / WT), WT, 0, s>>>(vec.p, x.p, p, gamma, q); KERNEL_CHECK();
                        compact(nullptr, (int)p, supp[cur_supp].p);
                    }
                } else {
                    q.thresh = alpha_f * q.pen_f;                                  // active_set_update(): all Scalar
                    q.denom = (float)(1.0 + (double)q.pen_f * (1.0 - (double)alpha_f));
                    if (nnz > 0) {
                        wide_tmp_kernel<<<zblocks, WT, 0, s>>>(Ax.p, z.p, y.p, frho, gamma, 1, n, tmp.p); KERNEL_CHECK();
                        if (fused_ok) { launch_fused(std::min(sms_n, nnz), supp[cur_supp].p, nnz, q, nullptr); fused_step = true; }
                        else {
                            const int blocks = std::min((nnz + 7) / 8, sm_count() * 8);
                            wide_active_kernel<<<blocks, WT, 0, s>>>(X, ldx, n, tmp.p, supp[cur_supp].p, nnz, x.p, q); KERNEL_CHECK();
                        }
                        compact(supp[cur_supp].p, nnz, supp[cur_supp ^ 1].p);
                        cur_supp ^= 1;
                    }
                }
                iter_counter++;
                step_kind = regular ? 1 : 2;
                // the new support size stays on the device: a regular step can grow the support to anything up to p,
                // an active-set step can only shrink it
                nnz_bound = regular ? (int)std::min<i64>(p, 2147483647LL) : nnz;
            }
            // ---------------- z step, residual, dual update ----------------
            const int chunks_bound = wide_chunks(nnz_bound, max_chunks);
            if (fused_step) {
                // (the partial A x of the fused step's logical CTAs are already in `part`)
                wide_zstep_kernel<<<zblocks, WT, 0, s>>>(part.p, nnz_dev.p, max_chunks, n, ydat.p, frho, (float)(-1 - rho), Ax.p, z.p, y.p, psums.p, nullptr, nnz_before, sms_n);
                KERNEL_CHECK();
            } else {
            if (chunks_bound > 0) {
                wide_ax_kernel<<<dim3((unsigned)zblocks, (unsigned)chunks_bound), WT, 0, s>>>(X, ldx, n, supp[cur_supp].p, nnz_dev.p, max_chunks, x.p, part.p);
                KERNEL_CHECK();
            }
            wide_zstep_kernel<<<zblocks, WT, 0, s>>>(part.p, nnz_dev.p, max_chunks, n, ydat.p, frho, (float)(-1 - rho), Ax.p, z.p, y.p, psums.p); KERNEL_CHECK();
            }
            wide_finish_sums_kernel<<<1, 192, 0, s>>>(psums.p, zblocks, nnz_dev.p, sums6.p); KERNEL_CHECK();
            double h[6];
            CUDA_CHECK(cudaMemcpyAsync(h, sums6.p, sizeof h, cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaStreamSynchronize(s));
            const double resid_dual = rho * (double)sqrt_sprad * (double)std::sqrt((float)h[0]);
            const double resid_primal = (double)std::sqrt((float)h[1]);
            sAx2 = h[2]; sz2 = h[3]; sy2 = h[4];
            nnz = (int)h[5];                                               // exact support size after this iteration's x step
            // (a screened regular step reads the fp16 copy of all columns and the float32 data of its candidates)
            const double xbytes = step_kind == 1 ? (screen ? 2.0 * (double)ldh * (double)p + 4.0 * (double)n * (double)last_ncand : 4.0 * (double)n * (double)p)
                                : step_kind == 2 ? 4.0 * (double)n * (double)nnz_before : 0.0;
            work_bytes += xbytes + 4.0 * (double)n * (double)nnz + 48.0 * (double)n;
            if (step_kind == 1) work_regular += 1; else if (step_kind == 2) work_active += 1;
            if (tracing && i < tr.cap) {
                double* row = tr.buf + 5 * (size_t)i;
                row[0] = eps_primal; row[1] = resid_primal; row[2] = eps_dual; row[3] = resid_dual; row[4] = rho;
                if (tr.nrows) *tr.nrows = i + 1;
            }
            if (resid_primal < eps_primal && resid_dual < eps_dual) break;
            if (i > 3) balance_rho(rho, resid_primal, eps_primal, resid_dual, eps_dual);
            i++;
        }
        out->niter[k] = i + 1;
        out->lambda[k] = lam[k];
        // solution of this lambda: support indices + values (supp[cur_supp] lists exactly the non-zeros of x)
        idx_all[k].resize(nnz); val_all[k].resize(nnz);
        if (nnz > 0) {
            wide_gather_kernel<<<(nnz + WT - 1) / WT, WT, 0, s>>>(x.p, supp[cur_supp].p, nnz, vec.p); KERNEL_CHECK();
            CUDA_CHECK(cudaMemcpyAsync(idx_all[k].data(), supp[cur_supp].p, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaMemcpyAsync(val_all[k].data(), vec.p, (size_t)nnz * sizeof(float), cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaStreamSynchronize(s));      // vec / supp are rewritten by the next lambda's first step
        }
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
    T.iterate = tm.stop();
    g_last_work[0] = work_bytes; g_last_work[1] = work_regular; g_last_work[2] = work_active;
    g_last_work[3] = screened_steps > 0 ? screened_cand / screened_steps : 0;      // mean number of columns a screened step evaluates exactly

    tm.start();
    out->nlambda = nl;
    finish_sparse_path(idx_all, val_all, nl, p, flag, meanX, scaleX, meanY, scaleY, out);
    T.finish = tm.stop();
    T.total = wall_now() - t_begin;
    out->rho = rho; out->eig = sprad; out->lambda0 = lambda0; out->t = T;
}

}  // namespace b200
