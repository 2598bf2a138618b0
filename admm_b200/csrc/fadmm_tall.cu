// fadmm_tall.cu -- the per-iteration hot path of the lasso / elastic-net solver for n > p,
// as ONE persistent cooperative kernel that runs the whole lambda path on the device.
//
// Reference being replaced (all in /root/reference/src):
//   FADMMBase::solve / update_x / update_z / update_y / converged      FADMMBase.h:185-265
//   ADMMLassoTall::next_x / next_z / next_residual / soft_threshold     ADMMLassoTall.h:55-95
//   ADMMLassoTall::compute_eps_* / compute_resid_* / diff_squared_norm  ADMMLassoTall.h:102-161
//   ADMMLassoTall::init_warm                                            ADMMLassoTall.h:219-230
//   ADMMEnetTall::enet                                                  ADMMEnet.h:24-45
//   the lambda loop of admm_lasso()                                     Lasso.cpp:97-124
//
// Data flow of one iteration (grid = one CTA per SM, 512 threads, state vectors in L2):
//   [A] x_own = Kinv[own rows, :] * rhs        rhs (p floats) lives in shared memory; each CTA
//       streams its contiguous row panel of Kinv = (X'X + rho I)^-1 from HBM exactly once with
//       128-bit L1-bypassing loads, 8 loads in flight per lane.  This is the only HBM traffic
//       of the iteration: 4 p^2 bytes.
//   [B] own rows: z = prox(x + adj_y/rho); r = x - z; y = adj_y + rho r; six partial sums
//       (|r|^2, |z-z_old|^2, |z-adj_z|^2, |x|^2, |z|^2, |y|^2): warp shuffle -> CTA -> global slot.
//   ---- one grid barrier (release/acquire counter) ----
//   [C] every CTA reduces the G x 6 partials in the same fixed order (bitwise identical
//       scalars everywhere), evaluates the stopping rule, the combined residual and the
//       restart test, then forms adj_z / adj_y for ALL p entries straight into the next rhs
//       in shared memory (4 L2-resident vectors), storing only its own rows of adj_*.
// z and y are triple-buffered and the partial sums double-buffered, which is what makes a
// single barrier per iteration race-free (a fast CTA can be at most one barrier ahead).
//
// Row-sharded runs (one process per GPU, TallPathArgs::nranks > 1): every rank holds the same K^-1 and
// owns the row range [p r / N, p (r + 1) / N); its CTAs stream only those rows (400 MB / N per iteration:
// L2-resident from N = 4 on).  The exchange of an iteration is fused into the kernel over NVLink peer
// memory: the new z / y entries of the own rows and the CTA's six partial sums are stored straight into
// every peer's state block (cudaIpc-mapped), and the grid barrier becomes a two-level barrier -- the local
// release / acquire counter, then one system-scope flag per source rank written by CTA 0 into each peer
// and polled by all CTAs.  Phase [C] reduces the N x G partials in rank-major order, so all ranks compute
// bit-identical scalars and iterates; no NCCL call and no host round trip inside the lambda path.
//
// Mixed precision follows the reference: vectors float, scalars double; double scalars are
// rounded to float where Eigen would do so (rho * r, adj_y / rho, (1+t) * z), and the prox
// compares/subtracts in double (ADMMLassoTall.h:64-67).  No FMA contraction is allowed on
// those expressions (explicit __fmul_rn/__fadd_rn), so one iteration differs from the CPU
// restatement only through the summation order of the norms and of the K^-1 product.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace b200 {

namespace {

constexpr int TP_THREADS = 512;
constexpr int TP_WARPS = TP_THREADS / 32;
constexpr int TP_SEG = 4;            // each row is split into 4 segments -> (row, segment) tasks
constexpr int NSUM = 6;
constexpr int PART_STRIDE = 8;

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// all CTAs of the (co-resident, cooperative) grid
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        red_release_add_u64(ctr, 1ULL);
        while (ld_acquire_u64(ctr) < target) { }
        __threadfence();
    }
    __syncthreads();
}
// All CTAs of all ranks, two levels.  After the CTA-wide sync ONE thread fences at system scope (which orders
// the peer stores of every thread of the CTA, exactly as in a cooperative grid sync) and arrives on the local
// counter; when the local grid is complete, CTA 0 publishes the barrier number to every peer (one release.sys
// store per peer) and every CTA waits until all peers have published theirs (acquire.sys).  Same-address
// atomics serialise at ~27 cycles each, so the arrivals stay local (G per counter) and only N - 1 words cross
// NVLink per rank and barrier.  A rank can be at most one barrier ahead of another.
// Bounded wait: a peer that never arrives (lost process) raises *abort instead of hanging the GPU.
__device__ __forceinline__ bool multi_rank_barrier(const TallPathArgs& a, unsigned long long nbar, int G, float* const* blocks,
                                                   volatile int* s_abort, long long* tp = nullptr)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        bool ok = true;
        const long long t0 = clock64();
        const long long limit = 12000000000LL;                         // ~6 s at 2 GHz
        __threadfence_system();
        if (tp) tp[0] = clock64();
        red_release_add_u64(a.barrier, 1ULL);
        while (ld_acquire_u64(a.barrier) < nbar * (unsigned long long)G) {
            if (*(volatile int*)a.abort_flag || clock64() - t0 > limit) { ok = false; break; }
        }
        if (tp) tp[1] = clock64();
        if (ok) {
            if (blockIdx.x == 0) {
                // ONE system fence, then relaxed stores: a release store per peer repeats the fence N - 1 times
                // (measured 3.2 us each -- the peer wait was 10 us at N = 4 and 22 us at N = 8)
                __threadfence_system();
                for (int k = 0; k < a.nranks; k++)
                    if (k != a.rank)
                        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(reinterpret_cast<unsigned long long*>(blocks[k] + a.off_flags) + a.rank), "l"(nbar) : "memory");
            }
            // poll all peers' flags together (independent relaxed loads: one L2 round trip per sweep, not one per
            // peer), then a single acquire fence
            const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(blocks[a.rank] + a.off_flags);
            for (;;) {
                unsigned long long lo = ~0ULL;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    unsigned long long v = ~0ULL;
                    if (k < a.nranks && k != a.rank) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + k) : "memory");
                    lo = v < lo ? v : lo;
                }
                if (lo >= nbar) break;
                if (*(volatile int*)a.abort_flag || clock64() - t0 > limit) { ok = false; break; }
            }
            __threadfence_system();
        }
        if (!ok) atomicExch(a.abort_flag, 1);
        *s_abort = ok ? 0 : 1;
    }
    __syncthreads();
    return *s_abort == 0;
}

struct ProxParams {
    int enet;
    double pen;          // lambda / rho
    float thresh, denom; // enet only
};
__device__ __forceinline__ float prox_apply(float v, const ProxParams& q)
{
    if (!q.enet) {
        if ((double)v > q.pen) return (float)((double)v - q.pen);
        if ((double)v < -q.pen) return (float)((double)v + q.pen);
        return 0.f;
    }
    if (v > q.thresh) return __fdiv_rn(__fsub_rn(v, q.thresh), q.denom);
    if (v < -q.thresh) return __fdiv_rn(__fadd_rn(v, q.thresh), q.denom);
    return 0.f;
}
__device__ __forceinline__ ProxParams make_prox(int enet, float lambda, double rho, double alpha_d)
{
    ProxParams q;
    q.enet = enet;
    q.pen = (double)lambda / rho;
    const float alpha = (float)alpha_d;
    q.thresh = (float)((double)alpha * q.pen);
    q.denom = (float)(1.0 + q.pen * (1.0 - (double)alpha));
    return q;
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(TP_THREADS, 1) tall_path_kernel(TallPathArgs a, int rows_per_cta, int ld)
{
    extern __shared__ __align__(16) float smem[];
    float* rhs = smem;                                   // ld floats (ld = p rounded up to 4)
    float* xpart = smem + ld;                            // rows_per_cta * TP_SEG
    __shared__ double s_sum[NSUM];
    __shared__ float s_red[TP_WARPS][NSUM];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int p = a.p;
    const int NR = a.nranks > 1 ? a.nranks : 1;
    const int R0 = NR > 1 ? (int)((long long)p * a.rank / NR) : 0, R1 = NR > 1 ? (int)((long long)p * (a.rank + 1) / NR) : p;
    const int r0 = min(R1, R0 + cta * rows_per_cta), r1 = min(R1, r0 + rows_per_cta);
    __shared__ int s_abort;
    unsigned long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // CTA 0 / thread 0: cycles in [A], [B], fence, local barrier, peer wait, [C] scalars, [C] total; iterations
    float* blocks[8];
#pragma unroll
    for (int k = 0; k < 8; k++) blocks[k] = a.peers[k];
    const int nrows = r1 - r0;
    const int ntasks = nrows * TP_SEG;
    const int segv = ((ld / 4) + TP_SEG - 1) / TP_SEG;   // float4 per segment
    const int nvec = ld / 4;

    auto zb = [&](int i) -> float* { return a.state + (size_t)i * ld; };          // triple-buffered z
    auto yb = [&](int i) -> float* { return a.state + (size_t)(3 + i) * ld; };    // triple-buffered y
    float* adj_z = a.state + 6 * (size_t)ld;
    float* adj_y = a.state + 7 * (size_t)ld;
    float* partials = a.state + 8 * (size_t)ld;          // [2][NR * G][PART_STRIDE]
    const size_t off_z = 0, off_y = 3 * (size_t)ld, off_part = 8 * (size_t)ld;
    const int NG = NR * G;

    const double rho = a.rho;
    const float frho = (float)rho;
    const double sqrt_p = sqrt((double)p);

    // scalars replicated (bitwise identically) in every thread of every CTA
    double sx2 = 0.0, sz2 = 0.0, sy2 = 0.0;              // |x|^2, |z|^2, |y|^2 of the current iterate
    double adj_a = 1.0, adj_c = 9999.0;
    int cur = 0;                                          // buffer holding the current z / y
    unsigned long long nbar = 0;
    unsigned git = 0;                                     // global iteration counter (buffer parity)

    for (int k = 0; k < a.nl; k++) {
        const float lambda = (float)a.lambdas[k];
        const ProxParams prox = make_prox(a.enet, lambda, rho, a.alpha);
        const bool tracing = (a.trace != nullptr) && (k == a.trace_lambda);

        // rhs = XY - adj_y + rho * adj_z from the stored extrapolation (cold start: zeros).  On a warm start this
        // is the rhs already in shared memory: a converged lambda leaves the loop before a new extrapolation is
        // formed, a lambda that ran out of iterations leaves right after phase [C] wrote the rhs of the new one.
        // It must NOT be rebuilt from global adj_*: after an exhausted lambda the other CTAs' rows of adj_* have
        // just been written with no grid barrier in between (and a sharded run stores its own rows only).
        if (k == 0)
        for (int v = tid; v < nvec; v += TP_THREADS) {
            const float4 xy = __ldg(reinterpret_cast<const float4*>(a.XY) + v);
            const float4 ay = ldcg4(adj_y + 4 * v), az = ldcg4(adj_z + 4 * v);
            float4 o;
            o.x = (float)((double)__fsub_rn(xy.x, ay.x) + rho * (double)az.x);
            o.y = (float)((double)__fsub_rn(xy.y, ay.y) + rho * (double)az.y);
            o.z = (float)((double)__fsub_rn(xy.z, ay.z) + rho * (double)az.z);
            o.w = (float)((double)__fsub_rn(xy.w, ay.w) + rho * (double)az.w);
            reinterpret_cast<float4*>(rhs)[v] = o;
        }
        __syncthreads();

        int niter = a.maxit + 1;
        for (int it = 0; it < a.maxit; it++, git++) {
            const int nxt = (cur + 1) % 3;
            const long long tA0 = clock64();
            // tolerances from the iterate before this step (FADMMBase.h:187-188)
            const double eps_primal = fmax((double)sqrtf((float)sx2), (double)sqrtf((float)sz2)) * a.eps_rel + sqrt_p * a.eps_abs;
            const double eps_dual = (double)sqrtf((float)sy2) * a.eps_rel + sqrt_p * a.eps_abs;

            // ---- [A] x_own = Kinv[own rows] * rhs -------------------------------------------------
            const bool rev = a.snake && (git & 1u);
            for (int t0 = warp; t0 < ntasks; t0 += TP_WARPS) {
                const int t = rev ? (ntasks - 1 - t0) : t0;
                const int row = t / TP_SEG, seg = t % TP_SEG;
                const float4* m = reinterpret_cast<const float4*>(a.Kinv + (size_t)(r0 + row) * ld);
                const float4* rv = reinterpret_cast<const float4*>(rhs);
                const int v0 = seg * segv, v1 = min(nvec, v0 + segv);
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                // 8 predicated 128-bit loads in flight per lane, also in the ragged last sweep
                for (int v = v0 + lane; v < v1; v += 8 * 32) {
                    float4 q[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int idx = v + u * 32;
                        q[u] = idx < v1 ? ld_stream_f4(m + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 8; u += 4) {
                        const int i0 = min(v + u * 32, v1 - 1), i1 = min(v + (u + 1) * 32, v1 - 1);
                        const int i2 = min(v + (u + 2) * 32, v1 - 1), i3 = min(v + (u + 3) * 32, v1 - 1);
                        const float4 b0 = rv[i0], b1 = rv[i1], b2 = rv[i2], b3 = rv[i3];
                        s0 = fmaf(q[u].x, b0.x, s0); s0 = fmaf(q[u].y, b0.y, s0); s0 = fmaf(q[u].z, b0.z, s0); s0 = fmaf(q[u].w, b0.w, s0);
                        s1 = fmaf(q[u + 1].x, b1.x, s1); s1 = fmaf(q[u + 1].y, b1.y, s1); s1 = fmaf(q[u + 1].z, b1.z, s1); s1 = fmaf(q[u + 1].w, b1.w, s1);
                        s2 = fmaf(q[u + 2].x, b2.x, s2); s2 = fmaf(q[u + 2].y, b2.y, s2); s2 = fmaf(q[u + 2].z, b2.z, s2); s2 = fmaf(q[u + 2].w, b2.w, s2);
                        s3 = fmaf(q[u + 3].x, b3.x, s3); s3 = fmaf(q[u + 3].y, b3.y, s3); s3 = fmaf(q[u + 3].z, b3.z, s3); s3 = fmaf(q[u + 3].w, b3.w, s3);
                    }
                }
                const float s = warp_sum((s0 + s1) + (s2 + s3));
                if (lane == 0) xpart[t] = s;
            }
            __syncthreads();

            // ---- [B] own rows: z, residual, y, partial sums -------------------------------------
            const long long tB0 = clock64();
            float ps[NSUM] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int r = tid; r < nrows; r += TP_THREADS) {
                const int i = r0 + r;
                const float xv = (xpart[r * TP_SEG] + xpart[r * TP_SEG + 1]) + (xpart[r * TP_SEG + 2] + xpart[r * TP_SEG + 3]);
                const float ay = __ldcg(adj_y + i), az = __ldcg(adj_z + i), zo = __ldcg(zb(cur) + i);
                const float v = __fadd_rn(xv, __fdiv_rn(ay, frho));
                const float zn = prox_apply(v, prox);
                const float res = __fsub_rn(xv, zn);
                const float yn = __fadd_rn(ay, __fmul_rn(frho, res));
                zb(nxt)[i] = zn;
                yb(nxt)[i] = yn;
                if (NR > 1) {
                    for (int kk = 0; kk < NR; kk++) {
                        if (kk == a.rank) continue;
                        blocks[kk][off_z + (size_t)nxt * ld + i] = zn;
                        blocks[kk][off_y + (size_t)nxt * ld + i] = yn;
                    }
                }
                const float d1 = zn - zo, d2 = zn - az;
                ps[0] += res * res; ps[1] += d1 * d1; ps[2] += d2 * d2;
                ps[3] += xv * xv;   ps[4] += zn * zn; ps[5] += yn * yn;
            }
#pragma unroll
            for (int q = 0; q < NSUM; q++) ps[q] = warp_sum(ps[q]);
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < NSUM; q++) s_red[warp][q] = ps[q];
            }
            __syncthreads();
            if (tid < NSUM) {
                float s = 0.f;
                for (int w = 0; w < TP_WARPS; w++) s += s_red[w][tid];
                const size_t slot = ((size_t)(git & 1u) * NG + (size_t)a.rank * (NR > 1 ? G : 0) + cta) * PART_STRIDE + tid;
                partials[slot] = s;
                for (int kk = 0; kk < NR; kk++)
                    if (NR > 1 && kk != a.rank) blocks[kk][off_part + slot] = s;
            }

            if (a.prof && cta == 0 && tid == 0) { const long long t = clock64(); pf[0] += (unsigned long long)(tB0 - tA0); pf[1] += (unsigned long long)(t - tB0); }
            nbar++;
            if (NR == 1) grid_barrier(a.barrier, nbar * (unsigned long long)G);
            else {
                long long tp[2] = {0, 0};
                const long long tb0 = clock64();
                if (!multi_rank_barrier(a, nbar, G, blocks, &s_abort, (a.prof && cta == 0) ? tp : nullptr)) return;
                if (a.prof && cta == 0 && tid == 0) {
                    const long long tb1 = clock64();
                    pf[2] += (unsigned long long)(tp[0] - tb0);      // CTA sync + system fence
                    pf[3] += (unsigned long long)(tp[1] - tp[0]);    // local grid barrier
                    pf[4] += (unsigned long long)(tb1 - tp[1]);      // flag publication + wait for the peers
                }
            }
            const long long tC0 = clock64();

            // ---- [C] global scalars, identical in every CTA ----------------------------------------
            if (warp < NSUM) {
                double s = 0.0;
                const float* src = partials + (size_t)(git & 1u) * NG * PART_STRIDE + warp;
                for (int c = lane; c < NG; c += 32) s += (double)__ldcg(src + (size_t)c * PART_STRIDE);
                s = warp_sum(s);
                if (lane == 0) s_sum[warp] = s;
            }
            __syncthreads();
            const double sum_r2 = s_sum[0], sum_dz2 = s_sum[1], sum_da2 = s_sum[2];
            sx2 = s_sum[3]; sz2 = s_sum[4]; sy2 = s_sum[5];
            const double resid_primal = (double)sqrtf((float)sum_r2);
            const double resid_dual = rho * sqrt((double)(float)sum_dz2);
            const int old = cur;
            cur = nxt;

            if (tracing && cta == 0 && tid == 0 && it < a.trace_cap) {
                double* row = a.trace + 5 * (size_t)it;
                row[0] = eps_primal; row[1] = resid_primal; row[2] = eps_dual; row[3] = resid_dual; row[4] = rho;
                *a.trace_rows = it + 1;
            }
            if (a.prof && cta == 0 && tid == 0) { pf[5] += (unsigned long long)(clock64() - tC0); pf[7] += 1ULL; }
            if (resid_primal < eps_primal && resid_dual < eps_dual) { niter = it + 1; git++; break; }

            const double old_c = adj_c;
            adj_c = rho * resid_primal * resid_primal + rho * (double)(float)sum_da2;
            bool accel;
            float c1 = 0.f, c2 = 0.f;
            if (adj_c < 0.999 * old_c) {
                const double old_a = adj_a;
                adj_a = 0.5 + 0.5 * sqrt(1.0 + 4.0 * old_a * old_a);
                const double ratio = (old_a - 1.0) / adj_a;
                c1 = (float)(1.0 + ratio); c2 = (float)ratio;
                accel = true;
            } else {
                adj_a = 1.0;
                adj_c = old_c / 0.999;
                accel = false;
            }
            // next rhs for all entries; own rows of adj_* stored for phase [B] / later warm starts
            const float* zn_ = zb(cur); const float* zo_ = zb(old);
            const float* yn_ = yb(cur); const float* yo_ = yb(old);
            for (int v = tid; v < nvec; v += TP_THREADS) {
                const float4 zo = ldcg4(zo_ + 4 * v), yo = ldcg4(yo_ + 4 * v);
                float4 az, ay;
                if (accel) {
                    const float4 zn = ldcg4(zn_ + 4 * v), yn = ldcg4(yn_ + 4 * v);
                    az.x = __fsub_rn(__fmul_rn(c1, zn.x), __fmul_rn(c2, zo.x));
                    az.y = __fsub_rn(__fmul_rn(c1, zn.y), __fmul_rn(c2, zo.y));
                    az.z = __fsub_rn(__fmul_rn(c1, zn.z), __fmul_rn(c2, zo.z));
                    az.w = __fsub_rn(__fmul_rn(c1, zn.w), __fmul_rn(c2, zo.w));
                    ay.x = __fsub_rn(__fmul_rn(c1, yn.x), __fmul_rn(c2, yo.x));
                    ay.y = __fsub_rn(__fmul_rn(c1, yn.y), __fmul_rn(c2, yo.y));
                    ay.z = __fsub_rn(__fmul_rn(c1, yn.z), __fmul_rn(c2, yo.z));
                    ay.w = __fsub_rn(__fmul_rn(c1, yn.w), __fmul_rn(c2, yo.w));
                } else { az = zo; ay = yo; }
                const float4 xy = __ldg(reinterpret_cast<const float4*>(a.XY) + v);
                float4 o;
                o.x = (float)((double)__fsub_rn(xy.x, ay.x) + rho * (double)az.x);
                o.y = (float)((double)__fsub_rn(xy.y, ay.y) + rho * (double)az.y);
                o.z = (float)((double)__fsub_rn(xy.z, ay.z) + rho * (double)az.z);
                o.w = (float)((double)__fsub_rn(xy.w, ay.w) + rho * (double)az.w);
                reinterpret_cast<float4*>(rhs)[v] = o;
                const int i = 4 * v;
                if (i + 3 >= r0 && i < r1) {
                    const float azv[4] = {az.x, az.y, az.z, az.w}, ayv[4] = {ay.x, ay.y, ay.z, ay.w};
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (i + e >= r0 && i + e < r1) { adj_z[i + e] = azv[e]; adj_y[i + e] = ayv[e]; }
                }
            }
            __syncthreads();
            if (a.prof && cta == 0 && tid == 0) pf[6] += (unsigned long long)(clock64() - tC0);
        }
        if (niter == a.maxit + 1) { /* loop ran out: git already advanced by the for-increment */ }

        // solution at this lambda = current z (own rows)
        for (int r = tid; r < nrows; r += TP_THREADS) {
            const float zv = __ldcg(zb(cur) + r0 + r);
            a.z_out[(size_t)k * p + r0 + r] = zv;
            for (int kk = 0; kk < NR; kk++)
                if (NR > 1 && kk != a.rank) blocks[kk][a.off_zout + (size_t)k * p + r0 + r] = zv;
        }
        if (cta == 0 && tid == 0) a.niter_out[k] = niter;
    }
    // the peers' rows of z_out have landed in this rank's block once everybody has passed this point
    if (NR > 1) { nbar++; multi_rank_barrier(a, nbar, G, blocks, &s_abort); }
    if (a.prof && cta == 0 && tid == 0)
        for (int q = 0; q < 8; q++) a.prof[q] = pf[q];
}

// ---------------------------------------------------------------------------------------------------------------
// Single-GPU form that reads ONE TRIANGLE of the symmetric K^-1 per iteration (2 p^2 bytes instead of 4 p^2).
//
// Row i of K^-1 is used up to the diagonal only:  K(i, j), j <= i, serves the dot product of row i (x_i += K(i, j) rhs_j)
// AND, by symmetry, row j's product (x_j += K(i, j) rhs_i, j < i).  The second contribution belongs to a row that another
// CTA owns, so phase [A] ends with per-CTA partial vectors and a grid barrier:
//   [A] CTA c streams the lower-triangle part of its rows.  Rows are folded -- CTA c owns the "top" rows [t0, t1) and the
//       "bottom" rows [p - t1, p - t0): every folded pair has p + 1 entries, so all CTAs stream the same number of bytes.
//       Thread t owns the columns 4 (t + 512 s) .. + 3 of every 2048-column stripe s: the column sums (the transposed
//       contributions) accumulate in the thread's own slots of a shared-memory vector -- no atomics, fixed order -- and the
//       row sums are reduced across the warp eight rows at a time by a 9-shuffle butterfly, then across the 16 warps.
//       The CTA stores its partial vector part[c][0 .. p) to global memory (L2-resident: G x p floats).
//   ---- grid barrier ----
//   [B] own rows: x_i = (row sum) + sum_c part[c][i] in CTA order (eight threads per row, fixed tree), then the prox, the
//       residual, the dual update and the six partial norms exactly as in tall_path_kernel.
//   ---- grid barrier ----   [C] as in tall_path_kernel.
// Everything is summed in a fixed order: runs are bit-reproducible; against tall_path_kernel only the summation order of
// the K^-1 product differs.  Needs 2 p floats of shared memory (rhs + column sums): p <= ~25 000; row-sharded runs keep
// tall_path_kernel (their exchange is per row block).
// Measured at p = 1e4 (profiles/r2k_*, r2l_*): DRAM reads per iteration 328 MB -> 134 MB (half the matrix, and the snake
// sweep keeps a larger share of it in the 126 MB L2), iteration 66.3 -> 58.6 us.  The gain is smaller than the byte count
// because neither kernel is HBM-bound at the clock the iteration phase runs at in a whole fit (~1.4 GHz right after the
// power-capped Gram): an SM takes in ~34 B / cycle from L2 (full-row kernel: 400 MB in 84.5 k cycles = 32 B / cycle / SM),
// and this sweep spends two FMAs per element plus the row-sum butterflies.  Two ring-buffered variants of phase [A]
// (per-thread cp.async, TMA bulk copies with full / empty mbarriers; git history) were slower: 63 and 75 us; so was a
// warp-tile sweep (warp = 128-column tiles, chunks of eight rows, column sums in registers): 86.8 k cycles for phase [A]
// against 65.0 k (profiles/r2t_*).
// ---------------------------------------------------------------------------------------------------------------
constexpr int TRI_ROWS = 8;                 // rows per group (loads in flight per thread)
constexpr int TRI_STRIPE = TP_THREADS * 4;  // columns per stripe

// v[0..7] per lane -> the lanes with (lane & 3) == 0 return the warp total of row ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)
__device__ __forceinline__ float butterfly8(float (&v)[TRI_ROWS], int lane)
{
    const unsigned full = 0xffffffffu;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float send = h16 ? v[k] : v[k + 4];
        const float keep = h16 ? v[k + 4] : v[k];
        v[k] = keep + __shfl_xor_sync(full, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const float send = h8 ? v[k] : v[k + 2];
        const float keep = h8 ? v[k + 2] : v[k];
        v[k] = keep + __shfl_xor_sync(full, send, 8);
    }
    {
        const float send = h4 ? v[0] : v[1];
        const float keep = h4 ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(full, send, 4);
    }
    v[0] += __shfl_xor_sync(full, v[0], 2);
    v[0] += __shfl_xor_sync(full, v[0], 1);
    return v[0];
}

__global__ void __launch_bounds__(TP_THREADS, 1) tall_path_tri_kernel(TallPathArgs a, int vrows_per_cta, int ld)
{
    extern __shared__ __align__(16) float smem[];
    float* rhs = smem;                                   // ld floats
    float* acc = smem + ld;                              // ld floats: column sums of this CTA's rows
    float* s_dot = acc + ld;                             // [TP_WARPS][2 * vpad]: per-warp row sums
    const int vpad = (vrows_per_cta + TRI_ROWS - 1) / TRI_ROWS * TRI_ROWS;
    float* xown = s_dot + TP_WARPS * 2 * vpad;           // [2 * vpad]: row sums of the own rows
    __shared__ double s_sum[NSUM];
    __shared__ float s_red[TP_WARPS][NSUM];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int p = a.p;
    const int V = (p + 1) / 2;                           // folded rows: v <-> rows v and p - 1 - v
    const int t0 = min(V, cta * vrows_per_cta), t1 = min(V, t0 + vrows_per_cta);
    const int b0 = max(p - t1, t1), b1 = p - t0;         // bottom rows (the middle row of an odd p belongs to the top block)
    const int nT = t1 - t0, nB = max(0, b1 - b0);
    const int nown = nT + nB;
    const int nvec = ld / 4;
    auto own_row = [&](int r) -> int { return r < nT ? t0 + r : b0 + (r - nT); };
    auto is_own = [&](int i) -> bool { return (i >= t0 && i < t1) || (i >= b0 && i < b1); };

    auto zb = [&](int i) -> float* { return a.state + (size_t)i * ld; };
    auto yb = [&](int i) -> float* { return a.state + (size_t)(3 + i) * ld; };
    float* adj_z = a.state + 6 * (size_t)ld;
    float* adj_y = a.state + 7 * (size_t)ld;
    float* partials = a.state + 8 * (size_t)ld;          // [2][G][PART_STRIDE]
    float* part = a.tri_part;                            // [G][ld]

    const double rho = a.rho;
    const float frho = (float)rho;
    const double sqrt_p = sqrt((double)p);
    double sx2 = 0.0, sz2 = 0.0, sy2 = 0.0;
    double adj_a = 1.0, adj_c = 9999.0;
    int cur = 0;
    unsigned long long nbar = 0;
    unsigned git = 0;
    unsigned long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    for (int k = 0; k < a.nl; k++) {
        const float lambda = (float)a.lambdas[k];
        const ProxParams prox = make_prox(a.enet, lambda, rho, a.alpha);
        const bool tracing = (a.trace != nullptr) && (k == a.trace_lambda);

        if (k == 0)
        for (int v = tid; v < nvec; v += TP_THREADS) {
            const float4 xy = __ldg(reinterpret_cast<const float4*>(a.XY) + v);
            const float4 ay = ldcg4(adj_y + 4 * v), az = ldcg4(adj_z + 4 * v);
            float4 o;
            o.x = (float)((double)__fsub_rn(xy.x, ay.x) + rho * (double)az.x);
            o.y = (float)((double)__fsub_rn(xy.y, ay.y) + rho * (double)az.y);
            o.z = (float)((double)__fsub_rn(xy.z, ay.z) + rho * (double)az.z);
            o.w = (float)((double)__fsub_rn(xy.w, ay.w) + rho * (double)az.w);
            reinterpret_cast<float4*>(rhs)[v] = o;
        }
        __syncthreads();

        int niter = a.maxit + 1;
        for (int it = 0; it < a.maxit; it++, git++) {
            const int nxt = (cur + 1) % 3;
            const long long tA0 = clock64();
            const double eps_primal = fmax((double)sqrtf((float)sx2), (double)sqrtf((float)sz2)) * a.eps_rel + sqrt_p * a.eps_abs;
            const double eps_dual = (double)sqrtf((float)sy2) * a.eps_rel + sqrt_p * a.eps_abs;

            // ---- [A] lower-triangle sweep of the own rows --------------------------------------------
            for (int v = tid; v < nvec; v += TP_THREADS) reinterpret_cast<float4*>(acc)[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool rev = a.snake && (git & 1u);
            const int ngT = (nT + TRI_ROWS - 1) / TRI_ROWS, ngB = (nB + TRI_ROWS - 1) / TRI_ROWS;
            for (int gg = 0; gg < ngT + ngB; gg++) {
                const int g = rev ? (ngT + ngB - 1 - gg) : gg;
                const bool bottom = g >= ngT;
                const int g0 = (bottom ? g - ngT : g) * TRI_ROWS;              // first row of the group within its block
                const int blk0 = bottom ? b0 : t0, blkn = bottom ? nB : nT;
                const int i0 = blk0 + g0;                                      // smallest row of the group
                const int nr = min(TRI_ROWS, blkn - g0);
                const int imax = i0 + nr - 1;
                float d[TRI_ROWS], ri[TRI_ROWS];
#pragma unroll
                for (int r = 0; r < TRI_ROWS; r++) { d[r] = 0.f; ri[r] = r < nr ? rhs[i0 + r] : 0.f; }
                const float* kbase = a.Kinv + (size_t)i0 * ld;
                for (int c4 = tid; 4 * c4 <= imax; c4 += TP_THREADS) {
                    const int j0 = 4 * c4;
                    float4 q[TRI_ROWS];
#pragma unroll
                    for (int r = 0; r < TRI_ROWS; r++)
                        q[r] = (r < nr && j0 <= i0 + r) ? ld_stream_f4(reinterpret_cast<const float4*>(kbase + (size_t)r * ld) + c4)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 rj = reinterpret_cast<const float4*>(rhs)[c4];
                    float4 av = reinterpret_cast<float4*>(acc)[c4];
                    if (j0 + 3 < i0) {
                        // all four columns lie strictly below the diagonal for every row of the group
#pragma unroll
                        for (int r = 0; r < TRI_ROWS; r++) {
                            d[r] = fmaf(q[r].x, rj.x, d[r]); d[r] = fmaf(q[r].y, rj.y, d[r]);
                            d[r] = fmaf(q[r].z, rj.z, d[r]); d[r] = fmaf(q[r].w, rj.w, d[r]);
                            av.x = fmaf(q[r].x, ri[r], av.x); av.y = fmaf(q[r].y, ri[r], av.y);
                            av.z = fmaf(q[r].z, ri[r], av.z); av.w = fmaf(q[r].w, ri[r], av.w);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < TRI_ROWS; r++) {
                            const int i = i0 + r;                              // rows beyond nr were loaded as zeros
                            const float qx = j0 <= i ? q[r].x : 0.f, qy = j0 + 1 <= i ? q[r].y : 0.f;
                            const float qz = j0 + 2 <= i ? q[r].z : 0.f, qw = j0 + 3 <= i ? q[r].w : 0.f;
                            d[r] = fmaf(qx, rj.x, d[r]); d[r] = fmaf(qy, rj.y, d[r]);
                            d[r] = fmaf(qz, rj.z, d[r]); d[r] = fmaf(qw, rj.w, d[r]);
                            av.x = fmaf(j0 < i ? qx : 0.f, ri[r], av.x); av.y = fmaf(j0 + 1 < i ? qy : 0.f, ri[r], av.y);
                            av.z = fmaf(j0 + 2 < i ? qz : 0.f, ri[r], av.z); av.w = fmaf(j0 + 3 < i ? qw : 0.f, ri[r], av.w);
                        }
                    }
                    reinterpret_cast<float4*>(acc)[c4] = av;
                }
                const float tot = butterfly8(d, lane);
                if ((lane & 3) == 0) {
                    const int r = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                    s_dot[warp * 2 * vpad + (bottom ? vpad : 0) + g0 + r] = tot;
                }
            }
            // the CTA's partial column sums -> global (every thread stores the slots it owns)
            {
                float4* dst = reinterpret_cast<float4*>(part + (size_t)cta * ld);
                for (int v = tid; v < nvec; v += TP_THREADS) __stcg(dst + v, reinterpret_cast<const float4*>(acc)[v]);
            }
            __syncthreads();
            for (int r = tid; r < 2 * vpad; r += TP_THREADS) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < TP_WARPS; w++) s += s_dot[w * 2 * vpad + r];
                xown[r] = s;
            }
            const long long tB0 = clock64();
            nbar++;
            grid_barrier(a.barrier, nbar * (unsigned long long)G);

            // ---- [B] own rows: x, z, residual, y, partial sums ------------------------------------
            float ps[NSUM] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int rb = 0; rb < nown; rb += TP_THREADS / 8) {
                const int r = rb + (tid >> 3), sub = tid & 7;
                const bool act = r < nown;
                const int i = act ? own_row(r) : 0;
                float s = 0.f;
                if (act) {
                    // all loads first (G <= 148 -> at most 19 per thread), then the sum in CTA order
                    float pv[19];
#pragma unroll
                    for (int q = 0; q < 19; q++) { const int c = sub + 8 * q; pv[q] = c < G ? __ldcg(part + (size_t)c * ld + i) : 0.f; }
#pragma unroll
                    for (int q = 0; q < 19; q++) s += pv[q];
                    for (int c = sub + 8 * 19; c < G; c += 8) s += __ldcg(part + (size_t)c * ld + i);
                }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                if (act && sub == 0) {
                    const float xv = __fadd_rn(xown[r < nT ? r : vpad + (r - nT)], s);
                    const float ay = __ldcg(adj_y + i), az = __ldcg(adj_z + i), zo = __ldcg(zb(cur) + i);
                    const float v = __fadd_rn(xv, __fdiv_rn(ay, frho));
                    const float zn = prox_apply(v, prox);
                    const float res = __fsub_rn(xv, zn);
                    const float yn = __fadd_rn(ay, __fmul_rn(frho, res));
                    zb(nxt)[i] = zn;
                    yb(nxt)[i] = yn;
                    const float d1 = zn - zo, d2 = zn - az;
                    ps[0] += res * res; ps[1] += d1 * d1; ps[2] += d2 * d2;
                    ps[3] += xv * xv;   ps[4] += zn * zn; ps[5] += yn * yn;
                }
            }
#pragma unroll
            for (int q = 0; q < NSUM; q++) ps[q] = warp_sum(ps[q]);
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < NSUM; q++) s_red[warp][q] = ps[q];
            }
            __syncthreads();
            if (tid < NSUM) {
                float s = 0.f;
                for (int w = 0; w < TP_WARPS; w++) s += s_red[w][tid];
                partials[((size_t)(git & 1u) * G + cta) * PART_STRIDE + tid] = s;
            }
            if (a.prof && cta == 0 && tid == 0) { const long long t = clock64(); pf[0] += (unsigned long long)(tB0 - tA0); pf[1] += (unsigned long long)(t - tB0); }
            nbar++;
            grid_barrier(a.barrier, nbar * (unsigned long long)G);
            const long long tC0 = clock64();

            // ---- [C] global scalars, identical in every CTA ----------------------------------------
            if (warp < NSUM) {
                double s = 0.0;
                const float* src = partials + (size_t)(git & 1u) * G * PART_STRIDE + warp;
                for (int c = lane; c < G; c += 32) s += (double)__ldcg(src + (size_t)c * PART_STRIDE);
                s = warp_sum(s);
                if (lane == 0) s_sum[warp] = s;
            }
            __syncthreads();
            const double sum_r2 = s_sum[0], sum_dz2 = s_sum[1], sum_da2 = s_sum[2];
            sx2 = s_sum[3]; sz2 = s_sum[4]; sy2 = s_sum[5];
            const double resid_primal = (double)sqrtf((float)sum_r2);
            const double resid_dual = rho * sqrt((double)(float)sum_dz2);
            const int old = cur;
            cur = nxt;

            if (tracing && cta == 0 && tid == 0 && it < a.trace_cap) {
                double* row = a.trace + 5 * (size_t)it;
                row[0] = eps_primal; row[1] = resid_primal; row[2] = eps_dual; row[3] = resid_dual; row[4] = rho;
                *a.trace_rows = it + 1;
            }
            if (a.prof && cta == 0 && tid == 0) { pf[5] += (unsigned long long)(clock64() - tC0); pf[7] += 1ULL; }
            if (resid_primal < eps_primal && resid_dual < eps_dual) { niter = it + 1; git++; break; }

            const double old_c = adj_c;
            adj_c = rho * resid_primal * resid_primal + rho * (double)(float)sum_da2;
            bool accel;
            float c1 = 0.f, c2 = 0.f;
            if (adj_c < 0.999 * old_c) {
                const double old_a = adj_a;
                adj_a = 0.5 + 0.5 * sqrt(1.0 + 4.0 * old_a * old_a);
                const double ratio = (old_a - 1.0) / adj_a;
                c1 = (float)(1.0 + ratio); c2 = (float)ratio;
                accel = true;
            } else {
                adj_a = 1.0;
                adj_c = old_c / 0.999;
                accel = false;
            }
            const float* zn_ = zb(cur); const float* zo_ = zb(old);
            const float* yn_ = yb(cur); const float* yo_ = yb(old);
            for (int v = tid; v < nvec; v += TP_THREADS) {
                const float4 zo = ldcg4(zo_ + 4 * v), yo = ldcg4(yo_ + 4 * v);
                float4 az, ay;
                if (accel) {
                    const float4 zn = ldcg4(zn_ + 4 * v), yn = ldcg4(yn_ + 4 * v);
                    az.x = __fsub_rn(__fmul_rn(c1, zn.x), __fmul_rn(c2, zo.x));
                    az.y = __fsub_rn(__fmul_rn(c1, zn.y), __fmul_rn(c2, zo.y));
                    az.z = __fsub_rn(__fmul_rn(c1, zn.z), __fmul_rn(c2, zo.z));
                    az.w = __fsub_rn(__fmul_rn(c1, zn.w), __fmul_rn(c2, zo.w));
                    ay.x = __fsub_rn(__fmul_rn(c1, yn.x), __fmul_rn(c2, yo.x));
                    ay.y = __fsub_rn(__fmul_rn(c1, yn.y), __fmul_rn(c2, yo.y));
                    ay.z = __fsub_rn(__fmul_rn(c1, yn.z), __fmul_rn(c2, yo.z));
                    ay.w = __fsub_rn(__fmul_rn(c1, yn.w), __fmul_rn(c2, yo.w));
                } else { az = zo; ay = yo; }
                const float4 xy = __ldg(reinterpret_cast<const float4*>(a.XY) + v);
                float4 o;
                o.x = (float)((double)__fsub_rn(xy.x, ay.x) + rho * (double)az.x);
                o.y = (float)((double)__fsub_rn(xy.y, ay.y) + rho * (double)az.y);
                o.z = (float)((double)__fsub_rn(xy.z, ay.z) + rho * (double)az.z);
                o.w = (float)((double)__fsub_rn(xy.w, ay.w) + rho * (double)az.w);
                reinterpret_cast<float4*>(rhs)[v] = o;
                const int i = 4 * v;
                const float azv[4] = {az.x, az.y, az.z, az.w}, ayv[4] = {ay.x, ay.y, ay.z, ay.w};
#pragma unroll
                for (int e = 0; e < 4; e++)
                    if (is_own(i + e)) { adj_z[i + e] = azv[e]; adj_y[i + e] = ayv[e]; }
            }
            __syncthreads();
            if (a.prof && cta == 0 && tid == 0) pf[6] += (unsigned long long)(clock64() - tC0);
        }

        for (int r = tid; r < nown; r += TP_THREADS) {
            const int i = own_row(r);
            a.z_out[(size_t)k * p + i] = __ldcg(zb(cur) + i);
        }
        if (cta == 0 && tid == 0) a.niter_out[k] = niter;
    }
    if (a.prof && cta == 0 && tid == 0)
        for (int q = 0; q < 8; q++) a.prof[q] = pf[q];
}

// stand-alone fused pass over long vectors (HBM-bound when len >> L2)
constexpr int ZU_THREADS = 256;
__global__ void __launch_bounds__(ZU_THREADS) fused_zu_kernel(const float* __restrict__ x, const float* __restrict__ adj_y,
                                                              const float* __restrict__ old_z, const float* __restrict__ adj_z,
                                                              float* __restrict__ z, float* __restrict__ y, i64 len,
                                                              float lambda, double rho, int enet, double alpha, float* __restrict__ part)
{
    const ProxParams prox = make_prox(enet, lambda, rho, alpha);
    const float frho = (float)rho;
    float ps[NSUM] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const i64 nvec = len / 4;
    const i64 stride = (i64)gridDim.x * ZU_THREADS;
    auto one = [&](float xv, float ay, float zo, float az, float& zn, float& yn) {
        const float v = __fadd_rn(xv, __fdiv_rn(ay, frho));
        zn = prox_apply(v, prox);
        const float res = __fsub_rn(xv, zn);
        yn = __fadd_rn(ay, __fmul_rn(frho, res));
        const float d1 = zn - zo, d2 = zn - az;
        ps[0] += res * res; ps[1] += d1 * d1; ps[2] += d2 * d2;
        ps[3] += xv * xv;   ps[4] += zn * zn; ps[5] += yn * yn;
    };
    for (i64 v = (i64)blockIdx.x * ZU_THREADS + threadIdx.x; v < nvec; v += stride) {
        const float4 xv = ld_stream_f4(reinterpret_cast<const float4*>(x) + v);
        const float4 ay = ld_stream_f4(reinterpret_cast<const float4*>(adj_y) + v);
        const float4 zo = ld_stream_f4(reinterpret_cast<const float4*>(old_z) + v);
        const float4 az = ld_stream_f4(reinterpret_cast<const float4*>(adj_z) + v);
        float4 zn, yn;
        one(xv.x, ay.x, zo.x, az.x, zn.x, yn.x);
        one(xv.y, ay.y, zo.y, az.y, zn.y, yn.y);
        one(xv.z, ay.z, zo.z, az.z, zn.z, yn.z);
        one(xv.w, ay.w, zo.w, az.w, zn.w, yn.w);
        __stcs(reinterpret_cast<float4*>(z) + v, zn);
        __stcs(reinterpret_cast<float4*>(y) + v, yn);
    }
    if (blockIdx.x == 0) {
        for (i64 i = nvec * 4 + threadIdx.x; i < len; i += ZU_THREADS) {
            float zn, yn;
            one(x[i], adj_y[i], old_z[i], adj_z[i], zn, yn);
            z[i] = zn; y[i] = yn;
        }
    }
    __shared__ float s_red[ZU_THREADS / 32][NSUM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NSUM; q++) ps[q] = warp_sum(ps[q]);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NSUM; q++) s_red[warp][q] = ps[q];
    }
    __syncthreads();
    if (threadIdx.x < NSUM) {
        float s = 0.f;
        for (int w = 0; w < ZU_THREADS / 32; w++) s += s_red[w][threadIdx.x];
        part[(size_t)blockIdx.x * NSUM + threadIdx.x] = s;
    }
}
__global__ void fused_zu_finish_kernel(const float* __restrict__ part, int nblocks, double* __restrict__ sums)
{
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q >= NSUM) return;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += (double)part[(size_t)b * NSUM + q];
    s = warp_sum(s);
    if (lane == 0) sums[q] = s;
}

}  // namespace

size_t tall_state_floats(int p)
{
    const size_t ld = ((size_t)p + 3) & ~(size_t)3;
    return 8 * ld + 2 * (size_t)2048 * PART_STRIDE;          // partial slots: up to 8 ranks x 148 CTAs, double-buffered
}

// grid of the one-triangle kernel: folded rows per CTA and the grid size (0: the shape does not take this path)
static int tri_grid(int p, int sms, int* vrows_per_cta, size_t* smem_bytes)
{
    const int ld = (p + 3) & ~3;
    const int V = (p + 1) / 2;
    int G = std::min(sms, std::max(1, (V + 3) / 4));
    const int vr = (V + G - 1) / G;
    G = (V + vr - 1) / vr;
    const int vpad = (vr + TRI_ROWS - 1) / TRI_ROWS * TRI_ROWS;
    *vrows_per_cta = vr;
    *smem_bytes = sizeof(float) * (2 * (size_t)ld + (size_t)(TP_WARPS + 1) * 2 * vpad);
    return (*smem_bytes <= 200 * 1024) ? G : 0;
}
// host-only replay of the one-triangle kernel's row assignment (b200admm_k_tri_plan): rows[4 c .. 4 c + 3] = t0, t1, b0, b1 of CTA c
int tall_tri_plan(int p, int sms, int* rows, int cap, long long* smem_bytes)
{
    int vr; size_t sm;
    const int G = tri_grid(p, sms, &vr, &sm);
    if (smem_bytes) *smem_bytes = (long long)sm;
    if (G <= 0) return 0;
    if (rows) {
        if (cap < G) return -1;
        const int V = (p + 1) / 2;
        for (int c = 0; c < G; c++) {
            const int t0 = std::min(V, c * vr), t1 = std::min(V, t0 + vr);
            rows[4 * c] = t0; rows[4 * c + 1] = t1; rows[4 * c + 2] = std::max(p - t1, t1); rows[4 * c + 3] = p - t0;
        }
    }
    return G;
}

size_t tall_tri_part_floats(int p)
{
    int vr; size_t sm;
    const int G = tri_grid(p, sm_count(), &vr, &sm);
    return G > 0 ? (size_t)G * (size_t)((p + 3) & ~3) : 0;
}

int launch_tall_path(cudaStream_t s, const TallPathArgs& a)
{
    const int p = a.p;
    const int ld = (p + 3) & ~3;
    const int sms = sm_count();
    if (a.tri_part != nullptr && a.nranks <= 1) {
        int vr; size_t smem;
        const int G = tri_grid(p, sms, &vr, &smem);
        if (G > 0) {
            static size_t tri_smem_set = 0;
            if (smem > tri_smem_set) {
                CUDA_CHECK(cudaFuncSetAttribute(tall_path_tri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                tri_smem_set = smem;
            }
            int occ = 0;
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tall_path_tri_kernel, TP_THREADS, smem));
            if (occ < 1) throw CudaError("tall path (triangle) kernel does not fit on an SM");
            TallPathArgs args = a;
            int vri = vr, ldi = ld;
            void* params[] = { (void*)&args, (void*)&vri, (void*)&ldi };
            CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)tall_path_tri_kernel, dim3(G), dim3(TP_THREADS), params, smem, s));
            ++g_launch_count;
            return G;
        }
    }
    // enough rows per CTA to amortise the barrier; never more CTAs than SMs (co-residency)
    // rows of this rank (the kernel uses the same split); every rank must run the SAME G for the partial slots
    const int NR = a.nranks > 1 ? a.nranks : 1;
    if (NR > 8) throw ArgError("tall path: at most 8 ranks share one lambda path");
    int own = 0;
    for (int r = 0; r < NR; r++) own = std::max(own, (int)((long long)p * (r + 1) / NR) - (int)((long long)p * r / NR));
    int G = std::min(sms, std::max(1, (own + 7) / 8));
    int rows_per_cta = (own + G - 1) / G;
    rows_per_cta = (rows_per_cta + 3) & ~3;
    G = std::min(G, (own + rows_per_cta - 1) / rows_per_cta);
    if ((size_t)NR * G > 2048) throw ArgError("tall path: too many partial-sum slots");
    const size_t smem = sizeof(float) * ((size_t)ld + (size_t)rows_per_cta * TP_SEG);
    if (smem > 200 * 1024) throw ArgError("tall path: p too large for the shared-memory resident rhs (p <= 50000)");
    static size_t smem_set = 0;
    if (smem > smem_set) {
        CUDA_CHECK(cudaFuncSetAttribute(tall_path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    int occ = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tall_path_kernel, TP_THREADS, smem));
    if (occ < 1) throw CudaError("tall path kernel does not fit on an SM");
    TallPathArgs args = a;
    int rpc = rows_per_cta, ldi = ld;
    void* params[] = { (void*)&args, (void*)&rpc, (void*)&ldi };
    CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)tall_path_kernel, dim3(G), dim3(TP_THREADS), params, smem, s));
    ++g_launch_count;
    return G;
}

int fused_zu_blocks() { return sm_count() * 8; }

void fused_zu_pass(cudaStream_t s, const float* x, const float* adj_y, const float* old_z, const float* adj_z,
                   float* z, float* y, i64 len, double lambda, double rho, int enet, double alpha,
                   double* sums_dev, float* partial_work)
{
    const int nb = fused_zu_blocks();
    fused_zu_kernel<<<nb, ZU_THREADS, 0, s>>>(x, adj_y, old_z, adj_z, z, y, len, (float)lambda, rho, enet, alpha, partial_work);
    KERNEL_CHECK();
    fused_zu_finish_kernel<<<1, 32 * NSUM, 0, s>>>(partial_work, nb, sums_dev);
    KERNEL_CHECK();
}

}  // namespace b200
