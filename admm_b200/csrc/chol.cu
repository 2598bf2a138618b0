// chol.cu -- blocked Cholesky, triangular inverse and SPD inverse on the device (f32 / f64).
//
// Replaces the one-time `Eigen::LLT::compute` of the reference
// (/root/reference/src/ADMMLassoTall.h:204-205, src/PADMMLasso.h:59-60, src/ADMMLAD.h:189,
// src/ADMMBP.h:169) and turns the per-iteration `solver.solve(rhs)` (src/ADMMLassoTall.h:79,
// src/PADMMLasso.h:25, src/ADMMLAD.h:76) into a bandwidth-bound product with the explicit
// inverse: rho never changes on the tall / consensus paths (update_rho is a no-op,
// src/ADMMLassoTall.h:97), so  K^-1 = (X'X + rho I)^-1  is formed once and every iteration is
// x = K^-1 rhs -- the same 4 p^2 bytes as the two triangular solves, but with no dependency
// chain, which is what lets the loop run at HBM speed.
//
//   chol_lower        right-looking, NB = 128: diagonal block factored in shared memory by one
//                     CTA (which also emits the block's inverse), panel = A21 * inv(L11)',
//                     trailing update A22 -= L21 L21' through gemm<T>.
//   tri_inverse_lower W = L^-1 block column by block column (right to left).
//   gram_of_lower     K^-1 = W' W (full symmetric storage).
#include "common.cuh"
#include "kernels.h"
#include <cstring>
#include <cstdlib>
#include <type_traits>

namespace b200 {

namespace {

constexpr int NB = 128;
constexpr int DIAG_THREADS = 512;

// Factor the kb x kb diagonal block at A (lda) in shared memory; write L11 back (lower part)
// and inv(L11) (lower triangular, zeros above) to Dinv (kb x kb, ld = NB).
// Right-looking, all 512 threads on the rank-1 update (2-D mapping: no integer division, no idle
// upper half); the block inverse is then built column by column -- thread c runs the forward
// substitution L w = e_c with its column of the inverse held in shared memory (float) or, when two
// NB x NB tiles do not fit (double), in the output buffer.
template <class T>
__global__ void __launch_bounds__(DIAG_THREADS) chol_diag_kernel(T* __restrict__ A, i64 lda, int kb, int k0, T* __restrict__ Dinv, int* info, int winv_in_smem,
                                                                 int write_upper)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);            // S[c * LDS + r], column-major, LDS = NB + 1
    constexpr int LDS = NB + 1;
    T* Wm = S + NB * LDS;                              // inverse, same layout (only if winv_in_smem)
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;            // 32 x 16

    for (int c = ty; c < kb; c += DIAG_THREADS / 32)
        for (int r = tx; r < kb; r += 32)
            S[c * LDS + r] = (r >= c) ? A[(i64)r + (i64)c * lda] : T(0);
    __syncthreads();

    for (int j = 0; j < kb; j++) {
        if (tid == 0) {
            T d = S[j * LDS + j];
            if (!(d > T(0))) { if (*info == 0) *info = k0 + j + 1; d = T(1); }
            S[j * LDS + j] = sqrt(d);
        }
        __syncthreads();
        const T piv = S[j * LDS + j];
        for (int r = j + 1 + tid; r < kb; r += DIAG_THREADS) S[j * LDS + r] = S[j * LDS + r] / piv;
        __syncthreads();
        // S(r,c) -= S(r,j) * S(c,j) for j < c <= r < kb
        for (int c = j + 1 + ty; c < kb; c += DIAG_THREADS / 32) {
            const T lc = S[j * LDS + c];
            for (int r = c + tx; r < kb; r += 32) S[c * LDS + r] -= S[j * LDS + r] * lc;
        }
        __syncthreads();
    }

    for (int c = ty; c < kb; c += DIAG_THREADS / 32)
        for (int r = c + tx; r < kb; r += 32) A[(i64)r + (i64)c * lda] = S[c * LDS + r];
    // upper form (A = U'U, U = L'): the factor also goes into the upper triangle of the block
    if (write_upper)
        for (int r = ty; r < kb; r += DIAG_THREADS / 32)
            for (int c = tx; c < r; c += 32) A[(i64)c + (i64)r * lda] = S[c * LDS + r];

    // inverse of the lower-triangular block
    if (winv_in_smem) {
        for (int c = tid; c < kb; c += DIAG_THREADS) {
            T* w = Wm + c * LDS;
            w[c] = T(1) / S[c * LDS + c];
            for (int r = c + 1; r < kb; r++) {
                T acc0 = T(0), acc1 = T(0);
                int k = c;
                for (; k + 1 < r; k += 2) { acc0 += S[k * LDS + r] * w[k]; acc1 += S[(k + 1) * LDS + r] * w[k + 1]; }
                if (k < r) acc0 += S[k * LDS + r] * w[k];
                w[r] = -(acc0 + acc1) / S[r * LDS + r];
            }
        }
        __syncthreads();
        for (int c = ty; c < NB; c += DIAG_THREADS / 32)
            for (int r = tx; r < NB; r += 32)
                Dinv[(i64)c * NB + r] = (c < kb && r < kb && r >= c) ? Wm[c * LDS + r] : T(0);
    } else {
        for (int c = tid; c < kb; c += DIAG_THREADS) {
            T* w = Dinv + (i64)c * NB;
            for (int r = 0; r < c; r++) w[r] = T(0);
            w[c] = T(1) / S[c * LDS + c];
            for (int r = c + 1; r < kb; r++) {
                T acc = T(0);
                for (int k = c; k < r; k++) acc += S[k * LDS + r] * w[k];
                w[r] = -acc / S[r * LDS + r];
            }
            for (int r = kb; r < NB; r++) w[r] = T(0);
        }
        for (int c = kb + tid; c < NB; c += DIAG_THREADS)
            for (int r = 0; r < NB; r++) Dinv[(i64)c * NB + r] = T(0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// float32 diagonal block, fast form: the same result as chol_diag_kernel (L11 in place, inv(L11) to Dinv) in ~1/8 of
// the time.  chol_diag_kernel walks 128 columns with three CTA-wide barriers each and then lets one thread per column
// run an 8000-term serial substitution; here the block is cut into four 32-column panels:
//   (1) warp 0 factors the 32 x 32 diagonal block in registers (lane = row, columns exchanged by shuffle),
//   (2) warp 0 inverts it (lane = column of the inverse, the factor read as shared-memory broadcasts),
//   (3) all warps form the panel below,  L21 = A21 inv(L11)',  as a small product,
//   (4) all warps apply the rank-32 update to the rest of the block,
// and (5) the off-diagonal 32 x 32 blocks of the inverse follow by block distance d = 1, 2, 3:
//       W(i, j) = -W(i, i) * sum_{k = j}^{i - 1} L(i, k) W(k, j),  all blocks of one distance in parallel.
// Blocks with kb < 128 are padded with the identity, so every case runs the same four panels.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PB = 32;
constexpr int LDF = 132;
constexpr size_t DIAG_FAST_SMEM = sizeof(float) * (2 * NB * LDF + 3 * PB * (PB + 4));

// One column step of the in-register 32 x 32 factorisation, instantiated for every column index so that each
// register-array access has a compile-time index (a runtime-indexed `a[j]` in a partially unrolled loop turns into a
// 32-way select chain: measured 19 000 instructions per panel instead of 1 300).
template <int J> struct CholRow {
    static __device__ __forceinline__ void run(float (&a)[PB], float& d, int lane, int* info, int c0, int kb, int k0)
    {
        const unsigned full = 0xffffffffu;
        if (!(d > 0.f)) {
            if (lane == 0 && c0 + J < kb && *info == 0) *info = k0 + c0 + J + 1;
            d = 1.f;
        }
        float ri = rsqrtf(d);
        ri = ri * (1.5f - 0.5f * d * ri * ri);                                  // one Newton step: full float accuracy
        const float lj = lane > J ? a[J] * ri : (lane == J ? d * ri : 0.f);      // L(lane, J)
        a[J] = lj;
        if (J + 1 < PB) {
            // the next pivot first, so that the rest of this column's update fills the latency of its rsqrt
            const float l1 = __shfl_sync(full, lj, (J + 1) & (PB - 1));
            if (J + 1 <= lane) a[(J + 1) & (PB - 1)] -= lj * l1;
            d = __shfl_sync(full, a[(J + 1) & (PB - 1)], (J + 1) & (PB - 1));
        }
#pragma unroll
        for (int c = J + 2; c < PB; c++) {
            const float lc = __shfl_sync(full, lj, c);                          // L(c, J)
            if (c <= lane) a[c] -= lj * lc;
        }
        CholRow<J + 1>::run(a, d, lane, info, c0, kb, k0);
    }
};
template <> struct CholRow<PB> {
    static __device__ __forceinline__ void run(float (&)[PB], float&, int, int*, int, int, int) {}
};

__global__ void __launch_bounds__(DIAG_THREADS) chol_diag_fast_kernel(float* __restrict__ A, i64 lda, int kb, int k0, float* __restrict__ Dinv,
                                                                      int* info, int write_upper)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* S = reinterpret_cast<float*>(smem_raw);        // S[c * LDF + r]: lower triangle of the block, becomes L
    float* W = S + NB * LDF;                               // same layout: inv(L), lower
    float* Tp = W + NB * LDF;                              // [3][PB][TPL] scratch of step (5)
    constexpr int TPL = PB + 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;

    for (int idx = tid; idx < NB * NB; idx += DIAG_THREADS) {
        const int c = idx >> 7, r = idx & 127;
        float v = (r == c) ? 1.f : 0.f;
        if (r >= c && r < kb && c < kb) v = A[(i64)r + (i64)c * lda];
        S[c * LDF + r] = v;
        W[c * LDF + r] = 0.f;
    }
    __syncthreads();

    for (int q = 0; q < NB / PB; q++) {
        const int c0 = q * PB;
        if (warp == 0) {
            // (1) factor the diagonal block in registers: lane holds row `lane`, a[c] = D(lane, c) for c <= lane.
            // The next pivot (column j + 1, then its reciprocal square root) is started before the rest of the rank-1
            // update so that the update's shuffles fill the latency of the pivot chain.
            float a[PB];
#pragma unroll
            for (int c = 0; c < PB; c++) a[c] = (c <= lane) ? S[(c0 + c) * LDF + c0 + lane] : 0.f;
            float d = __shfl_sync(full, a[0], 0);
            CholRow<0>::run(a, d, lane, info, c0, kb, k0);
#pragma unroll
            for (int c = 0; c < PB; c++) if (c <= lane) S[(c0 + c) * LDF + c0 + lane] = a[c];
            // (2) inverse of the diagonal block, row-wise in registers: t[c] accumulates sum_{k' < k} L(lane, k') Winv(k', c);
            // at step k lane k finalises its row Winv(k, c) = (delta_kc - t[c]) / L(k, k) and broadcasts it.
            float t[PB];
#pragma unroll
            for (int c = 0; c < PB; c++) t[c] = 0.f;
            float dl = 0.f;                                        // L(lane, lane), picked out with static indices
#pragma unroll
            for (int c = 0; c < PB; c++) if (c == lane) dl = a[c];
            const float rdl = 1.f / dl;
#pragma unroll
            for (int k = 0; k < PB; k++) {
#pragma unroll
                for (int c = 0; c <= k; c++) {
                    // row k of the inverse, held by lane k
                    const float mine = ((c == k ? 1.f : 0.f) - t[c]) * rdl;
                    const float wkc = __shfl_sync(full, mine, k);                       // Winv(k, c)
                    if (lane == k) t[c] = wkc;                                          // lane k keeps its final row in t
                    else if (lane > k) t[c] += a[k] * wkc;
                }
            }
            // lane r now holds row r of the inverse in t[0 .. r]
#pragma unroll
            for (int c = 0; c < PB; c++) if (c <= lane) W[(c0 + c) * LDF + c0 + lane] = t[c];
        }
        __syncthreads();
        const int base = c0 + PB, nr = NB - base;                                       // rows below the panel
        if (nr > 0) {
            // (3) L21 = A21 inv(L11)':  X(r, j) = sum_k A21(r, k) Winv(j, k)   (Winv(j, k) = 0 for k > j)
            // thread = one row, eight columns
            const int rr = tid % nr, jg = tid / nr;
            const bool act3 = jg < PB / 8;
            float x8[8];
#pragma unroll
            for (int u = 0; u < 8; u++) x8[u] = 0.f;
            if (act3) {
                const int r = base + rr, j0 = jg * 8;
#pragma unroll 4
                for (int k = 0; k < PB; k++) {
                    const float av = S[(c0 + k) * LDF + r];
                    const float4 w0 = *reinterpret_cast<const float4*>(W + (c0 + k) * LDF + c0 + j0);
                    const float4 w1 = *reinterpret_cast<const float4*>(W + (c0 + k) * LDF + c0 + j0 + 4);
                    x8[0] += av * w0.x; x8[1] += av * w0.y; x8[2] += av * w0.z; x8[3] += av * w0.w;
                    x8[4] += av * w1.x; x8[5] += av * w1.y; x8[6] += av * w1.z; x8[7] += av * w1.w;
                }
            }
            __syncthreads();
            if (act3) {
                const int r = base + rr, j0 = jg * 8;
#pragma unroll
                for (int u = 0; u < 8; u++) S[(c0 + j0 + u) * LDF + r] = x8[u];
            }
            __syncthreads();
            // (4) rank-32 update of the rest, 4 x 4 register tiles on and below the diagonal:
            //     S(r, c) -= sum_j L21(r, j) L21(c, j)
            const int nt = nr / 4;
            for (int idx = tid; idx < nt * nt; idx += DIAG_THREADS) {
                const int tr = idx % nt, tc = idx / nt;
                if (tc > tr) continue;
                const int r0 = base + 4 * tr, cc0 = base + 4 * tc;
                float acc[4][4];
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) acc[x][y] = 0.f;
#pragma unroll 4
                for (int j = 0; j < PB; j++) {
                    const float4 lr = *reinterpret_cast<const float4*>(S + (c0 + j) * LDF + r0);
                    const float4 lc = *reinterpret_cast<const float4*>(S + (c0 + j) * LDF + cc0);
                    const float rv[4] = {lr.x, lr.y, lr.z, lr.w}, cv[4] = {lc.x, lc.y, lc.z, lc.w};
#pragma unroll
                    for (int x = 0; x < 4; x++)
#pragma unroll
                        for (int y = 0; y < 4; y++) acc[x][y] += rv[x] * cv[y];       // acc[x][y] -> S(r0 + x, cc0 + y)
                }
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    float4 o = *reinterpret_cast<float4*>(S + (cc0 + y) * LDF + r0);
                    o.x -= acc[0][y]; o.y -= acc[1][y]; o.z -= acc[2][y]; o.w -= acc[3][y];
                    *reinterpret_cast<float4*>(S + (cc0 + y) * LDF + r0) = o;
                }
            }
            __syncthreads();
        }
    }

    // (5) off-diagonal blocks of the inverse, by block distance: 4 x 4 register tiles, 64 tiles per 32 x 32 block
    for (int d = 1; d < NB / PB; d++) {
        const int npair = NB / PB - d;
        const bool act = tid < npair * 64;
        const int pr = tid >> 6, e = tid & 63, r0 = (e & 7) * 4, q0 = (e >> 3) * 4;     // rows r0.., columns q0.. of block (ib, jb)
        const int jb = pr, ib = pr + d;
        if (act) {
            float acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) acc[x][y] = 0.f;
            for (int kk = PB * jb; kk < PB * ib; kk += 4) {
                float4 wv[4];
#pragma unroll
                for (int y = 0; y < 4; y++) wv[y] = *reinterpret_cast<const float4*>(W + (PB * jb + q0 + y) * LDF + kk);   // Winv(kk .. kk+3, col)
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const float4 lv = *reinterpret_cast<const float4*>(S + (kk + u) * LDF + PB * ib + r0);                 // L(rows, kk + u)
                    const float rv[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
                    for (int y = 0; y < 4; y++) {
                        const float wk = u == 0 ? wv[y].x : (u == 1 ? wv[y].y : (u == 2 ? wv[y].z : wv[y].w));
#pragma unroll
                        for (int x = 0; x < 4; x++) acc[x][y] += rv[x] * wk;
                    }
                }
            }
#pragma unroll
            for (int x = 0; x < 4; x++)
                *reinterpret_cast<float4*>(Tp + (pr * PB + r0 + x) * TPL + q0) = make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
        }
        __syncthreads();
        if (act) {
            float acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) acc[x][y] = 0.f;
#pragma unroll 4
            for (int k = 0; k < PB; k++) {
                const float4 wi = *reinterpret_cast<const float4*>(W + (PB * ib + k) * LDF + PB * ib + r0);               // Winv_ii(rows, k)
                const float4 tv = *reinterpret_cast<const float4*>(Tp + (pr * PB + k) * TPL + q0);
                const float rv[4] = {wi.x, wi.y, wi.z, wi.w}, cv[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
                for (int x = 0; x < 4; x++)
#pragma unroll
                    for (int y = 0; y < 4; y++) acc[x][y] += rv[x] * cv[y];
            }
#pragma unroll
            for (int y = 0; y < 4; y++)
                *reinterpret_cast<float4*>(W + (PB * jb + q0 + y) * LDF + PB * ib + r0) = make_float4(-acc[0][y], -acc[1][y], -acc[2][y], -acc[3][y]);
        }
        __syncthreads();
    }

    for (int idx = tid; idx < NB * NB; idx += DIAG_THREADS) {
        const int c = idx >> 7, r = idx & 127;
        const bool in = r >= c && r < kb && c < kb;
        if (in) A[(i64)r + (i64)c * lda] = S[c * LDF + r];
        Dinv[(i64)c * NB + r] = in ? W[c * LDF + r] : 0.f;
    }
    if (write_upper)
        for (int idx = tid; idx < NB * NB; idx += DIAG_THREADS) {
            const int r = idx >> 7, c = idx & 127;                 // element (c, r) of the upper triangle <- L(r, c)
            if (c < r && r < kb) A[(i64)c + (i64)r * lda] = S[c * LDF + r];
        }
}

// U12 <- inv(L11) A12 in place, for the kb x m block row A12 to the right of a factored diagonal block
// (Zi = inv(L11): kb x kb lower triangular, leading dimension NB).  One CTA per 64 columns: both operands are
// staged in shared memory, so the block can be overwritten where it stands (no scratch panel, no copy back).
constexpr int TP_COLS = 64;
constexpr int TP_LDB = TP_COLS + 4;
constexpr size_t TRSM_PANEL_SMEM = sizeof(float) * (NB * NB + NB * TP_LDB);
__global__ void __launch_bounds__(256) trsm_panel_kernel(const float* __restrict__ Zi, int kb, float* __restrict__ A12, i64 lda, i64 m)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Zs = reinterpret_cast<float*>(smem_raw);       // Zs[k * NB + i] = Zi(i, k)
    float* Bs = Zs + NB * NB;                              // Bs[k * TP_LDB + j] = A12(k, j0 + j)
    const int tid = threadIdx.x;
    const i64 j0 = (i64)blockIdx.x * TP_COLS;
    const int nc = (int)min((i64)TP_COLS, m - j0);
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int k = idx >> 7, i = idx & 127;
        Zs[idx] = (i < kb && k < kb) ? Zi[(i64)k * NB + i] : 0.f;
    }
    for (int idx = tid; idx < NB * TP_COLS; idx += 256) {
        const int j = idx >> 7, k = idx & 127;
        Bs[k * TP_LDB + j] = (k < kb && j < nc) ? A12[(i64)k + (j0 + j) * lda] : 0.f;
    }
    __syncthreads();
    // thread: rows i0 .. i0 + 7, columns c0 .. c0 + 3
    const int i0 = (tid & 15) * 8, c0 = (tid >> 4) * 4;
    float acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
    const int kmax = min(kb, i0 + 8);                      // Zi(i, k) = 0 for k > i
    for (int k = 0; k < kmax; k++) {
        const float4 z0 = *reinterpret_cast<const float4*>(Zs + k * NB + i0);
        const float4 z1 = *reinterpret_cast<const float4*>(Zs + k * NB + i0 + 4);
        const float4 bv = *reinterpret_cast<const float4*>(Bs + k * TP_LDB + c0);
        const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = fmaf(zz[a], bb[b], acc[a][b]);
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) Bs[(i0 + a) * TP_LDB + c0 + b] = acc[a][b];
    __syncthreads();
    for (int idx = tid; idx < NB * TP_COLS; idx += 256) {
        const int j = idx >> 7, k = idx & 127;
        if (k < kb && j < nc) A12[(i64)k + (j0 + j) * lda] = Bs[k * TP_LDB + j];
    }
}

template <class T>
__global__ void copy_block_kernel(const T* __restrict__ src, i64 lds, T* __restrict__ dst, i64 ldd, i64 rows, i64 cols, T scale)
{
    const i64 total = rows * cols;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const i64 r = idx % rows, c = idx / rows;
        dst[r + c * ldd] = scale * src[r + c * lds];
    }
}
template <class T>
void copy_block(cudaStream_t s, const T* src, i64 lds, T* dst, i64 ldd, i64 rows, i64 cols, T scale = T(1))
{
    if (rows <= 0 || cols <= 0) return;
    const i64 total = rows * cols;
    const unsigned grid = (unsigned)std::min<i64>((total + 255) / 256, (i64)sm_count() * 8);
    copy_block_kernel<T><<<grid, 256, 0, s>>>(src, lds, dst, ldd, rows, cols, scale);
    KERNEL_CHECK();
}

// single right-hand side substitution with the factor (one CTA; not on the hot path)
template <class T>
__global__ void __launch_bounds__(1024) chol_solve_vec_kernel(const T* __restrict__ L, int p, i64 lda, T* __restrict__ b)
{
    // forward: L w = b   (column-oriented)
    for (int j = 0; j < p; j++) {
        __syncthreads();
        const T wj = b[j] / L[(i64)j + (i64)j * lda];
        __syncthreads();
        if (threadIdx.x == 0) b[j] = wj;
        for (int r = j + 1 + threadIdx.x; r < p; r += blockDim.x) b[r] -= L[(i64)r + (i64)j * lda] * wj;
    }
    __syncthreads();
    // backward: L' x = w   (row j of L' is column j of L: dot product form)
    __shared__ T scratch[33];
    for (int j = p - 1; j >= 0; j--) {
        T acc = T(0);
        for (int r = j + 1 + threadIdx.x; r < p; r += blockDim.x) acc += L[(i64)r + (i64)j * lda] * b[r];
        acc = block_sum(acc, scratch);
        if (threadIdx.x == 0) b[j] = (b[j] - acc) / L[(i64)j + (i64)j * lda];
        __syncthreads();
    }
}

// out (cols x rows, ld ldo) = in' for in (rows x cols, ld ldi): 32 x 32 tiles through shared memory
template <class T>
__global__ void __launch_bounds__(256) transpose_block_kernel(const T* __restrict__ in, i64 ldi, i64 rows, i64 cols, T* __restrict__ out, i64 ldo)
{
    __shared__ T tile[32][33];
    const i64 r0 = (i64)blockIdx.x * 32, c0 = (i64)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int c = ty; c < 32; c += 8)
        tile[c][tx] = (r0 + tx < rows && c0 + c < cols) ? in[(r0 + tx) + (c0 + c) * ldi] : T(0);
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
        if (c0 + tx < cols && r0 + r < rows) out[(c0 + tx) + (r0 + r) * ldo] = tile[tx][r];
}

}  // namespace

// Two-level blocking.  Inner blocks of NB = 128 columns are factored as before (diagonal block in shared
// memory, panel = A21 inv(L11)'), but their rank-128 updates are confined to the current OUTER panel of
// OB = 1024 columns; the rest of the matrix then receives ONE rank-1024 update per outer panel,
// A22 -= L21 L21', which for float runs on the tensor cores: L21 is transposed into a K-major scratch
// array and handed to the tcgen05 3xTF32 CTA-pair Gram kernel in subtract mode (fp32-accurate; 97 % of
// the factorisation's flops at p = 1e4).  With single-level blocking the same flops were 79 rank-128
// CUDA-core updates at ~12 TFLOP/s.
constexpr int OB = 1024;

// work layout: [nblk * NB * NB] inverse diagonal blocks, [p * NB] panel scratch, [p * OB] transposed outer panel
template <class T> size_t chol_work(i64 p)
{
    const i64 nblk = (p + NB - 1) / NB;
    return (size_t)(nblk * NB * NB + p * NB + p * OB);
}
template size_t chol_work<float>(i64);
template size_t chol_work<double>(i64);
size_t tri_inverse_tmp(i64 p) { return (size_t)p * OB; }

template <class T>
void chol_lower(cudaStream_t s, T* A, i64 p, i64 lda, T* work, int* info_dev)
{
    const i64 nblk = (p + NB - 1) / NB;
    T* Dinv = work;
    T* panel = work + nblk * NB * NB;
    T* Lt = panel + p * NB;                                       // OB x m2, ld = OB
    const int winv_in_smem = sizeof(T) == 4 ? 1 : 0;             // two 128 x 129 tiles: 132 KB in float, too large in double
    const size_t smem = sizeof(T) * NB * (NB + 1) * (winv_in_smem ? 2 : 1);
    static bool attr_done_f = false, attr_done_d = false;
    bool& done = std::is_same<T, float>::value ? attr_done_f : attr_done_d;
    if (!done) {
        CUDA_CHECK(cudaFuncSetAttribute(chol_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done = true;
    }
    const char* tenv = getenv("B200ADMM_FACTOR_TENSOR");
    const bool tensor_ok = std::is_same<T, float>::value && !(tenv && !strcmp(tenv, "0"));
    CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int), s));
    for (i64 k0 = 0; k0 < p; k0 += OB) {
        const i64 kend = std::min<i64>(p, k0 + OB), ob = kend - k0;
        for (i64 j0 = k0; j0 < kend; j0 += NB) {
            const i64 b = j0 / NB;
            const int kb = (int)std::min<i64>(NB, p - j0);
            T* Akk = A + j0 + j0 * lda;
            chol_diag_kernel<T><<<1, DIAG_THREADS, smem, s>>>(Akk, lda, kb, (int)j0, Dinv + b * NB * NB, info_dev, winv_in_smem, 0);
            KERNEL_CHECK();
            const i64 m = p - j0 - kb;
            if (m <= 0) break;
            T* A21 = A + (j0 + kb) + j0 * lda;
            // L21 = A21 * inv(L11)'
            gemm<T>(s, false, true, m, kb, kb, T(1), A21, lda, Dinv + b * NB * NB, NB, T(0), panel, m, 0);
            copy_block<T>(s, panel, m, A21, lda, m, kb);
            // inside the outer panel: A(j0+kb.., j0+kb..kend) -= L21 L21(0..w,:)'   (lower part of the leading square)
            const i64 w = kend - (j0 + kb);
            if (w > 0)
                gemm<T>(s, false, true, m, w, kb, T(-1), panel, m, panel, m, T(1), A + (j0 + kb) + (j0 + kb) * lda, lda, GEMM_LOWER);
        }
        const i64 m2 = p - kend;
        if (m2 <= 0) break;
        const T* L21 = A + kend + k0 * lda;                        // m2 x ob
        T* A22 = A + kend + kend * lda;
        bool on_tensor = false;
        if (tensor_ok && m2 >= 256 && ob % 4 == 0) {
            dim3 tg((unsigned)((m2 + 31) / 32), (unsigned)((ob + 31) / 32));
            transpose_block_kernel<T><<<tg, 256, 0, s>>>(L21, lda, m2, ob, Lt, ob);
            KERNEL_CHECK();
            on_tensor = gram_tn_tensor_sub(s, reinterpret_cast<const float*>(Lt), ob, ob, m2, reinterpret_cast<float*>(A22), lda);
        }
        if (!on_tensor)
            gemm<T>(s, false, true, m2, m2, ob, T(-1), L21, lda, L21, lda, T(1), A22, lda, GEMM_LOWER);
    }
}
template void chol_lower<float>(cudaStream_t, float*, i64, i64, float*, int*);
template void chol_lower<double>(cudaStream_t, double*, i64, i64, double*, int*);

// W <- L^-1.  `work` is chol_lower's workspace (inverse diagonal blocks); tmp: tri_inverse_tmp(p) entries.
// Two-level as well: the OB x OB diagonal block of W is inverted with the NB recursion, then the whole
// block column below it follows from two products with N = OB columns,  T = L21 W11,  W21 = -W22 T,
// wide enough to fill the chip (the NB-wide products of the one-level recursion ran on 40-80 CTAs).
template <class T>
void tri_inverse_lower(cudaStream_t s, const T* L, i64 p, i64 lda, const T* work, T* W, i64 ldw, T* tmp)
{
    CUDA_CHECK(cudaMemsetAsync(W, 0, sizeof(T) * (size_t)ldw * (size_t)p, s));
    const i64 nouter = (p + OB - 1) / OB;
    for (i64 o = nouter - 1; o >= 0; o--) {
        const i64 k0 = o * OB, kend = std::min<i64>(p, k0 + OB), ob = kend - k0;
        // W11 = inv(L11), NB block columns right to left inside the outer block
        const i64 nin = (ob + NB - 1) / NB;
        for (i64 bi = nin - 1; bi >= 0; bi--) {
            const i64 j0 = k0 + bi * NB;
            const int kb = (int)std::min<i64>(NB, kend - j0);
            const T* Dinv = work + (j0 / NB) * NB * NB;
            copy_block<T>(s, Dinv, NB, W + j0 + j0 * ldw, ldw, kb, kb);
            const i64 m = kend - j0 - kb;
            if (m <= 0) continue;
            gemm<T>(s, false, false, m, kb, kb, T(1), L + (j0 + kb) + j0 * lda, lda, Dinv, NB, T(0), tmp, m, 0);
            gemm<T>(s, false, false, m, kb, m, T(-1), W + (j0 + kb) + (j0 + kb) * ldw, ldw, tmp, m, T(0),
                    W + (j0 + kb) + j0 * ldw, ldw, GEMM_A_LOWER_TRI);
        }
        const i64 m2 = p - kend;
        if (m2 <= 0) continue;
        // T = L21 * W11   (W11 lower triangular)
        gemm<T>(s, false, false, m2, ob, ob, T(1), L + kend + k0 * lda, lda, W + k0 + k0 * ldw, ldw, T(0), tmp, m2, GEMM_B_LOWER_TRI);
        // W21 = - W22 * T  (W22 lower triangular, already final)
        gemm<T>(s, false, false, m2, ob, m2, T(-1), W + kend + kend * ldw, ldw, tmp, m2, T(0), W + kend + k0 * ldw, ldw, GEMM_A_LOWER_TRI);
    }
}
template void tri_inverse_lower<float>(cudaStream_t, const float*, i64, i64, const float*, float*, i64, float*);
template void tri_inverse_lower<double>(cudaStream_t, const double*, i64, i64, const double*, double*, i64, double*);

template <class T>
void gram_of_lower(cudaStream_t s, const T* W, i64 p, i64 ldw, T* Kinv, i64 ldk)
{
    gemm<T>(s, true, false, p, p, p, T(1), W, ldw, W, ldw, T(0), Kinv, ldk,
            GEMM_LOWER | GEMM_MIRROR | GEMM_AT_LOWER_TRI | GEMM_B_LOWER_TRI);
}
// float: W'W is a Gram matrix of a K-major operand -- the tcgen05 3xTF32 kernel takes it as is
// (W is stored with explicit zeros above the diagonal)
template <>
void gram_of_lower<float>(cudaStream_t s, const float* W, i64 p, i64 ldw, float* Kinv, i64 ldk)
{
    if (p >= 512 && gram_tn_tensor(s, W, p, ldw, p, Kinv, ldk, 1)) return;
    gemm<float>(s, true, false, p, p, p, 1.f, W, ldw, W, ldw, 0.f, Kinv, ldk,
                GEMM_LOWER | GEMM_MIRROR | GEMM_AT_LOWER_TRI | GEMM_B_LOWER_TRI);
}
template void gram_of_lower<double>(cudaStream_t, const double*, i64, i64, double*, i64);

// ---------------------------------------------------------------------------------------------------------------
// float32, p >= 512: the whole of  a <- a^-1  with every O(p^3) product on the tensor cores (tn_pair_kernel, 3xTF32).
//
// The tensor-core operands are K-major, i.e. only products of the form C = A'B (A: K x M, B: K x N, column-major)
// run without a transposing copy.  The factorisation is therefore carried in its UPPER form on the column-major
// array, a = U'U (U = L'): the block row U12 of a panel is K x m column-major as it stands, so both trailing updates
//     a22 -= U12' U12      (rank 128 inside the current 1024-column outer panel, rank 1024 beyond it)
// are TN products.  (a is symmetric with both triangles stored, and every update rewrites whole 256 x 256 tiles on
// the diagonal, so the diagonal blocks the one-CTA kernel factors always find their lower triangle current.)
// The inverse Z = L^-1 (lower) is built block column by block column from the last outer panel to the first:
//     Z11 = inv(L11)  (128-column recursion on the CUDA cores, 3 % of the flops),
//     T   = L21 Z11   = U12' Z11,                      TN as it stands,
//     Z21 = -Z22 T    = -(Z22')' T,                    TN with the transposed copy Zt = Z' kept alongside
// (each finished block column is transposed into Zt once: p^2 / 2 elements over the whole run), the zero blocks of
// the triangular operands skipped by K range.  Finally  a = Z'Z  on the tiles below the diagonal, k >= max(I, J),
// and a mirror pass.  Total tensor flops p^3 against the ~3 p^3 of the version that treated L^-1 as dense.
// work: spd_tc_work_floats(p) floats; Z, Zt: p x ld each (Z is left holding L^-1).
// ---------------------------------------------------------------------------------------------------------------
static inline i64 round4(i64 v) { return (v + 3) & ~(i64)3; }
size_t spd_tc_work_floats(i64 p)
{
    const i64 nblk = (p + NB - 1) / NB, nouter = (p + OB - 1) / OB;
    return (size_t)(nblk * NB * NB + nouter * ((i64)OB * NB + (i64)OB * OB) + round4(p) * OB);
}
bool spd_inverse_tc_usable(i64 p, i64 ld)
{
    const char* fenv = getenv("B200ADMM_FACTOR");
    if (fenv && !strcmp(fenv, "legacy")) return false;
    const char* tenv = getenv("B200ADMM_FACTOR_TENSOR");
    if (tenv && !strcmp(tenv, "0")) return false;
    return p >= 512 && ld % 4 == 0 && (sm_count() % 2 == 0);
}
void spd_inverse_tc(cudaStream_t s, float* a, i64 p, i64 ld, float* Z, float* Zt, float* work, int* info_dev)
{
    typedef float T;
    const i64 nblk = (p + NB - 1) / NB;
    const i64 nouter = (p + OB - 1) / OB;
    T* Dinv = work;
    T* panels = Dinv + nblk * NB * NB;           // per outer block: OB x NB scratch of the 128-column recursion
    T* Lds = panels + nouter * (i64)OB * NB;     // per outer block: OB x OB transposed diagonal block (its lower triangle is L11)
    T* Tm = Lds + nouter * (i64)OB * OB;         // round4(m2) x OB
    const size_t smem = sizeof(T) * NB * (NB + 1) * 2;
    static bool attr_done = false;
    if (!attr_done) {
        CUDA_CHECK(cudaFuncSetAttribute(chol_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_CHECK(cudaFuncSetAttribute(chol_diag_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_FAST_SMEM));
        CUDA_CHECK(cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_PANEL_SMEM));
        attr_done = true;
    }
    const char* denv = getenv("B200ADMM_DIAG");
    const bool fast_diag = !(denv && !strcmp(denv, "legacy"));
    CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int), s));
    CUDA_CHECK(cudaMemsetAsync(Z, 0, sizeof(T) * (size_t)ld * (size_t)p, s));
    CUDA_CHECK(cudaMemsetAsync(Zt, 0, sizeof(T) * (size_t)ld * (size_t)p, s));

    // The inverse of each diagonal 1024 x 1024 block of L (a chain of ~14 small CUDA-core launches) needs nothing but
    // that block, so it is issued on side streams as soon as its outer panel is factored and runs on the SMs the
    // one-CTA / few-CTA kernels of the main chain leave idle; the main stream joins before the block products.
    constexpr int NSIDE = 2;
    static cudaStream_t side[NSIDE] = {nullptr, nullptr};
    static std::vector<cudaEvent_t> fork_ev;
    static cudaEvent_t join_ev[NSIDE] = {nullptr, nullptr};
    if (!side[0]) {
        for (int i = 0; i < NSIDE; i++) {
            CUDA_CHECK(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
            CUDA_CHECK(cudaEventCreateWithFlags(&join_ev[i], cudaEventDisableTiming));
        }
    }
    while ((i64)fork_ev.size() < nouter) {
        cudaEvent_t e;
        CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        fork_ev.push_back(e);
    }
    auto invert_diagonal_outer_block = [&](cudaStream_t q, i64 o) {
        const i64 k0 = o * OB, kend = std::min<i64>(p, k0 + OB), ob = kend - k0;
        T* Ld = Lds + o * (i64)OB * OB;
        T* panel = panels + o * (i64)OB * NB;
        dim3 tg((unsigned)((ob + 31) / 32), (unsigned)((ob + 31) / 32));
        transpose_block_kernel<T><<<tg, 256, 0, q>>>(a + k0 + k0 * ld, ld, ob, ob, Ld, OB);
        KERNEL_CHECK();
        const i64 nin = (ob + NB - 1) / NB;
        for (i64 bi = nin - 1; bi >= 0; bi--) {
            const i64 j0 = k0 + bi * NB, l0 = bi * NB;
            const int kb = (int)std::min<i64>(NB, kend - j0);
            const T* Di = Dinv + (j0 / NB) * NB * NB;
            copy_block<T>(q, Di, NB, Z + j0 + j0 * ld, ld, kb, kb);
            const i64 m = kend - j0 - kb;
            if (m <= 0) continue;
            gemm<T>(q, false, false, m, kb, kb, T(1), Ld + (l0 + kb) + l0 * OB, OB, Di, NB, T(0), panel, m, 0);
            gemm<T>(q, false, false, m, kb, m, T(-1), Z + (j0 + kb) + (j0 + kb) * ld, ld, panel, m, T(0),
                    Z + (j0 + kb) + j0 * ld, ld, GEMM_A_LOWER_TRI);
        }
    };

    // ---- a = U'U, upper triangle of a <- U -----------------------------------------------------------------
    for (i64 k0 = 0; k0 < p; k0 += OB) {
        const i64 kend = std::min<i64>(p, k0 + OB), ob = kend - k0;
        for (i64 j0 = k0; j0 < kend; j0 += NB) {
            const i64 b = j0 / NB;
            const int kb = (int)std::min<i64>(NB, p - j0);
            const i64 j1 = j0 + kb;
            if (fast_diag) chol_diag_fast_kernel<<<1, DIAG_THREADS, DIAG_FAST_SMEM, s>>>(a + j0 + j0 * ld, ld, kb, (int)j0, Dinv + b * NB * NB, info_dev, 1);
            else chol_diag_kernel<T><<<1, DIAG_THREADS, smem, s>>>(a + j0 + j0 * ld, ld, kb, (int)j0, Dinv + b * NB * NB, info_dev, 1, 1);
            KERNEL_CHECK();
            const i64 m = p - j1;
            if (m <= 0) break;
            T* U12 = a + j0 + j1 * ld;                               // kb x m, K-major as it stands
            // U12 = inv(L11) A12
            if (fast_diag) {
                trsm_panel_kernel<<<(unsigned)((m + TP_COLS - 1) / TP_COLS), 256, TRSM_PANEL_SMEM, s>>>(Dinv + b * NB * NB, kb, U12, ld, m);
                KERNEL_CHECK();
            } else {
                gemm<T>(s, false, false, kb, m, kb, T(1), Dinv + b * NB * NB, NB, U12, ld, T(0), Tm, NB, GEMM_A_LOWER_TRI);
                copy_block<T>(s, Tm, NB, U12, ld, kb, m);
            }
            const i64 w = kend - j1;                                 // rows left in this outer panel
            if (w > 0) {
                T* A22 = a + j1 + j1 * ld;
                const bool on_tensor = m >= 512 &&
                    gemm_tn_tensor(s, U12, ld, U12, ld, w, m, kb, A22, ld, TN_TILES_ALL, TN_K_ALL, TN_K_ALL, TN_SUB);
                if (!on_tensor) gemm<T>(s, true, false, w, m, kb, T(-1), U12, ld, U12, ld, T(1), A22, ld, 0);
            }
        }
        {
            const i64 o = k0 / OB;
            cudaStream_t q = side[o % NSIDE];
            CUDA_CHECK(cudaEventRecord(fork_ev[o], s));
            CUDA_CHECK(cudaStreamWaitEvent(q, fork_ev[o], 0));
            invert_diagonal_outer_block(q, o);
        }
        const i64 m2 = p - kend;
        if (m2 <= 0) break;
        T* U12o = a + k0 + kend * ld;                                // ob x m2
        T* A22o = a + kend + kend * ld;
        const bool on_tensor = m2 >= 256 &&
            gemm_tn_tensor(s, U12o, ld, U12o, ld, m2, m2, ob, A22o, ld, TN_TILES_UPPER, TN_K_ALL, TN_K_ALL, TN_SUB);
        if (!on_tensor) gemm<T>(s, true, false, m2, m2, ob, T(-1), U12o, ld, U12o, ld, T(1), A22o, ld, 0);
    }

    // ---- Z = L^-1 (lower), Zt = Z' (upper): the diagonal blocks are done (side streams), the rest follows -----------
    for (int i = 0; i < NSIDE && i < nouter; i++) {          // only streams that were forked (they are part of a capture)
        CUDA_CHECK(cudaEventRecord(join_ev[i], side[i]));
        CUDA_CHECK(cudaStreamWaitEvent(s, join_ev[i], 0));
    }
    for (i64 o = nouter - 1; o >= 0; o--) {
        const i64 k0 = o * OB, kend = std::min<i64>(p, k0 + OB), ob = kend - k0;
        const i64 m2 = p - kend;
        if (m2 > 0) {
            const i64 ldt = round4(m2);
            const T* U12o = a + k0 + kend * ld;                      // ob x m2  (= L21')
            const T* Z11 = Z + k0 + k0 * ld;
            // T = L21 Z11 = U12' Z11   (Z11 lower triangular: k from block J on)
            if (!gemm_tn_tensor(s, U12o, ld, Z11, ld, m2, ob, ob, Tm, ldt, TN_TILES_ALL, TN_K_J, TN_K_ALL, TN_STORE))
                gemm<T>(s, true, false, m2, ob, ob, T(1), U12o, ld, Z11, ld, T(0), Tm, ldt, GEMM_B_LOWER_TRI);
            // Z21 = -Z22 T = -(Zt22)' T   (Zt22 upper triangular: k up to the end of block I)
            if (!gemm_tn_tensor(s, Zt + kend + kend * ld, ld, Tm, ldt, m2, ob, m2, Z + kend + k0 * ld, ld, TN_TILES_ALL, TN_K_ALL, TN_K_I, TN_NEG))
                gemm<T>(s, false, false, m2, ob, m2, T(-1), Z + kend + kend * ld, ld, Tm, ldt, T(0), Z + kend + k0 * ld, ld, GEMM_A_LOWER_TRI);
        }
        if (o > 0) {
            // Zt[k0:kend, k0:p] = (Z[k0:p, k0:kend])'
            const i64 rows = p - k0;
            dim3 tg((unsigned)((rows + 31) / 32), (unsigned)((ob + 31) / 32));
            transpose_block_kernel<T><<<tg, 256, 0, s>>>(Z + k0 + k0 * ld, ld, rows, ob, Zt + k0 + k0 * ld, ld);
            KERNEL_CHECK();
        }
    }

    // ---- a = Z'Z (tiles on and below the diagonal, k >= max(I, J)), then the mirror image ----------------------
    if (gemm_tn_tensor(s, Z, ld, Z, ld, p, p, p, a, ld, TN_TILES_LOWER, TN_K_OUTER, TN_K_ALL, TN_STORE))
        mirror_lower_to_upper(s, a, p, ld);
    else
        gram_of_lower<float>(s, Z, p, ld, a, ld);
}

// micro-benchmark of the one-CTA diagonal-block kernels (tools/time_diag.py): mode 0 empty kernel with the same launch
// configuration, 1 chol_diag_kernel<float>, 2 chol_diag_fast_kernel; returns the mean milliseconds per launch
namespace { __global__ void __launch_bounds__(DIAG_THREADS) empty_smem_kernel(float* out) { extern __shared__ float sm[]; if (out == nullptr) sm[threadIdx.x] = 1.f; } }
double diag_kernel_bench(cudaStream_t s, int mode, int reps, float* A, i64 lda, float* Dinv, int* info)
{
    const size_t smem_old = sizeof(float) * NB * (NB + 1) * 2;
    CUDA_CHECK(cudaFuncSetAttribute(chol_diag_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_old));
    CUDA_CHECK(cudaFuncSetAttribute(chol_diag_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_FAST_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(empty_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_FAST_SMEM));
    EventTimer t(s);
    t.start();
    for (int r = 0; r < reps; r++) {
        if (mode == 0) empty_smem_kernel<<<1, DIAG_THREADS, DIAG_FAST_SMEM, s>>>(A);
        else if (mode == 1) chol_diag_kernel<float><<<1, DIAG_THREADS, smem_old, s>>>(A, lda, NB, 0, Dinv, info, 1, 1);
        else chol_diag_fast_kernel<<<1, DIAG_THREADS, DIAG_FAST_SMEM, s>>>(A, lda, NB, 0, Dinv, info, 1);
        KERNEL_CHECK();
    }
    return t.stop() * 1e3 / reps;
}

template <class T>
void chol_solve_vec(cudaStream_t s, const T* L, i64 p, i64 lda, T* b)
{
    chol_solve_vec_kernel<T><<<1, 1024, 0, s>>>(L, (int)p, lda, b);
    KERNEL_CHECK();
}
template void chol_solve_vec<float>(cudaStream_t, const float*, i64, i64, float*);
template void chol_solve_vec<double>(cudaStream_t, const double*, i64, i64, double*);

}  // namespace b200
