// gemm.cu -- general dense products on the CUDA cores (f32 and f64), column-major.
//
//   C (M x N) = alpha * op(A) (M x K) * op(B) (K x N) + beta * C
//
// This is the work-horse behind every one-time setup product that is not the big Gram
// matrix (which has its own tensor-core kernel in gram_tc.cu): the Cholesky panel and
// trailing updates, the triangular inverse, K^-1 = W'W, X X' for the wide / BP solvers,
// M = L^-1 A for BP (reference call sites: Linalg::cross_prod_lower / tcross_prod_lower,
// /root/reference/src/Linalg/BlasWrapper.h:73-154; dtrsm_ in src/ADMMLAD.h:198-199 and
// src/ADMMBP.h:181-182; Eigen::LLT::compute in src/ADMMLassoTall.h:204-205).
//
// Register-tiled (each thread owns a 2x2 arrangement of VxV sub-tiles so that shared-memory
// reads are 128-bit and conflict-free), global loads prefetched into registers one K-slab
// ahead.  Arbitrary M, N, K and leading dimensions; 64-bit addressing.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace b200 {

struct CfgLarge { static constexpr int BM = 128, BN = 128, BK = 16, V = 4; };   // 8 x 8 outputs per thread
struct CfgSmall { static constexpr int BM = 64,  BN = 64,  BK = 16, V = 2; };   // 4 x 4 outputs per thread
template <class T> struct GemmCfg;
template <> struct GemmCfg<float>  { typedef CfgLarge Large; typedef CfgSmall Small; };
template <> struct GemmCfg<double> { typedef CfgSmall Large; typedef CfgSmall Small; };

template <class T, class Cfg, bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_kernel(int M, int N, int K, T alpha, const T* __restrict__ A, i64 lda,
            const T* __restrict__ B, i64 ldb, T beta, T* __restrict__ C, i64 ldc, int mode)
{
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, V = Cfg::V;
    constexpr int PAD = 4;
    constexpr int LA = BM * BK / 256, LB = BN * BK / 256;   // elements each thread stages

    const int bi = blockIdx.x, bj = blockIdx.y;
    const int i0 = bi * BM, j0 = bj * BN;
    if ((mode & GEMM_LOWER) && j0 > i0 + BM - 1) return;     // tile strictly above the diagonal

    // K range: operands known to be lower-triangular let us skip structurally zero slabs
    int kbeg = 0, kend = K;
    if (mode & GEMM_A_LOWER_TRI) {          // op(A) = A, A lower triangular (M x K): A(i,k)=0 for k>i
        kend = min(K, i0 + BM);
    }
    if (mode & GEMM_AT_LOWER_TRI) {         // op(A) = A', A lower triangular (K x M): A(k,i)=0 for k<i
        kbeg = max(kbeg, (i0 / BK) * BK);
    }
    if (mode & GEMM_B_LOWER_TRI) {          // op(B) = B, B lower triangular (K x N): B(k,j)=0 for k<j
        kbeg = max(kbeg, (j0 / BK) * BK);
    }
    if (mode & GEMM_BT_LOWER_TRI) {         // op(B) = B', B lower triangular (N x K): B(j,k)=0 for k>j
        kend = min(kend, j0 + BN);
    }
    const bool mA = (mode & GEMM_A_LOWER_TRI) != 0, mAT = (mode & GEMM_AT_LOWER_TRI) != 0;
    const bool mB = (mode & GEMM_B_LOWER_TRI) != 0, mBT = (mode & GEMM_BT_LOWER_TRI) != 0;

    __shared__ __align__(16) T As[BK][BM + PAD];
    __shared__ __align__(16) T Bs[BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    T acc[2 * V][2 * V];
#pragma unroll
    for (int a = 0; a < 2 * V; a++)
#pragma unroll
        for (int b = 0; b < 2 * V; b++) acc[a][b] = T(0);

    T ra[LA], rb[LB];

    auto fetch = [&](int k0) {
#pragma unroll
        for (int l = 0; l < LA; l++) {
            const int idx = tid + l * 256;
            int i, k;
            if (!TA) { i = idx % BM; k = idx / BM; } else { k = idx % BK; i = idx / BK; }
            const int gi = i0 + i, gk = k0 + k;
            T v = T(0);
            if (gi < M && gk < kend) v = TA ? A[(i64)gk + (i64)gi * lda] : A[(i64)gi + (i64)gk * lda];
            if ((mA && gk > gi) || (mAT && gk < gi)) v = T(0);   // structural zeros of a triangular operand
            ra[l] = v;
        }
#pragma unroll
        for (int l = 0; l < LB; l++) {
            const int idx = tid + l * 256;
            int j, k;
            if (TB) { j = idx % BN; k = idx / BN; } else { k = idx % BK; j = idx / BK; }
            const int gj = j0 + j, gk = k0 + k;
            T v = T(0);
            if (gj < N && gk < kend) v = TB ? B[(i64)gj + (i64)gk * ldb] : B[(i64)gk + (i64)gj * ldb];
            if ((mB && gk < gj) || (mBT && gk > gj)) v = T(0);
            rb[l] = v;
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int l = 0; l < LA; l++) {
            const int idx = tid + l * 256;
            int i, k;
            if (!TA) { i = idx % BM; k = idx / BM; } else { k = idx % BK; i = idx / BK; }
            As[k][i] = ra[l];
        }
#pragma unroll
        for (int l = 0; l < LB; l++) {
            const int idx = tid + l * 256;
            int j, k;
            if (TB) { j = idx % BN; k = idx / BN; } else { k = idx % BK; j = idx / BK; }
            Bs[k][j] = rb[l];
        }
    };

    if (kbeg < kend) fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        __syncthreads();
        stage();
        __syncthreads();
        if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; k++) {
            T a[2 * V], b[2 * V];
#pragma unroll
            for (int v = 0; v < V; v++) {
                a[v] = As[k][ty * V + v];
                a[V + v] = As[k][BM / 2 + ty * V + v];
                b[v] = Bs[k][tx * V + v];
                b[V + v] = Bs[k][BN / 2 + tx * V + v];
            }
#pragma unroll
            for (int x = 0; x < 2 * V; x++)
#pragma unroll
                for (int y = 0; y < 2 * V; y++) acc[x][y] += a[x] * b[y];
        }
    }

    const bool lower = (mode & GEMM_LOWER) != 0, mirror = (mode & GEMM_MIRROR) != 0;
#pragma unroll
    for (int x = 0; x < 2 * V; x++) {
        const int gi = i0 + (x < V ? ty * V + x : BM / 2 + ty * V + (x - V));
        if (gi >= M) continue;
#pragma unroll
        for (int y = 0; y < 2 * V; y++) {
            const int gj = j0 + (y < V ? tx * V + y : BN / 2 + tx * V + (y - V));
            if (gj >= N) continue;
            if (lower && gj > gi) continue;
            T* c = C + (i64)gi + (i64)gj * ldc;
            T v = alpha * acc[x][y];
            if (beta != T(0)) v += beta * (*c);
            *c = v;
            if (mirror && gi != gj) C[(i64)gj + (i64)gi * ldc] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// float64 on the tensor cores: the same contract as gemm_kernel<double, ...> (all four transpositions, triangular K
// ranges and masks, lower / mirror stores) with the inner product issued as mma.sync.m8n8k4.f64 -- the LAD / BP Gram
// matrices and triangular solves (src/ADMMLAD.h:186-201, src/ADMMBP.h:167-182) are float64 in the reference and the
// parity bar there is 1e-9, so no split-precision scheme applies.  128 x 128 x 16 tiles, 16 warps of 32 x 32, operands
// staged K-major in shared memory (row stride 132 doubles: the 4 x 4 lane pattern of a fragment load touches 16
// distinct 8-byte banks), global loads prefetched into registers one K-slab ahead.  DFMA in IEEE order inside the
// instruction: results differ from the CUDA-core kernel by summation order only.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(512, 1)
gemm_f64_mma_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, i64 lda,
                    const double* __restrict__ B, i64 ldb, double beta, double* __restrict__ C, i64 ldc, int mode)
{
    constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;
    constexpr int NT = 512;                                   // 16 warps of 32 x 32: four warps per scheduler
    constexpr int LA = BM * BK / NT, LB = BN * BK / NT;

    const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
    if ((mode & GEMM_LOWER) && j0 > i0 + BM - 1) return;

    int kbeg = 0, kend = K;
    if (mode & GEMM_A_LOWER_TRI) kend = min(K, i0 + BM);
    if (mode & GEMM_AT_LOWER_TRI) kbeg = max(kbeg, (i0 / BK) * BK);
    if (mode & GEMM_B_LOWER_TRI) kbeg = max(kbeg, (j0 / BK) * BK);
    if (mode & GEMM_BT_LOWER_TRI) kend = min(kend, j0 + BN);
    const bool mA = (mode & GEMM_A_LOWER_TRI) != 0, mAT = (mode & GEMM_AT_LOWER_TRI) != 0;
    const bool mB = (mode & GEMM_B_LOWER_TRI) != 0, mBT = (mode & GEMM_BT_LOWER_TRI) != 0;

    __shared__ __align__(16) double As[BK][BM + PAD];
    __shared__ __align__(16) double Bs[BK][BN + PAD];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int fr = lane >> 2, fk = lane & 3;                 // fragment row / column index, k index

    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }

    double ra[LA], rb[LB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int l = 0; l < LA; l++) {
            const int idx = tid + l * NT;
            int i, k;
            if (!TA) { i = idx % BM; k = idx / BM; } else { k = idx % BK; i = idx / BK; }
            const int gi = i0 + i, gk = k0 + k;
            double v = 0.0;
            if (gi < M && gk < kend) v = TA ? A[(i64)gk + (i64)gi * lda] : A[(i64)gi + (i64)gk * lda];
            if ((mA && gk > gi) || (mAT && gk < gi)) v = 0.0;
            ra[l] = v;
        }
#pragma unroll
        for (int l = 0; l < LB; l++) {
            const int idx = tid + l * NT;
            int j, k;
            if (TB) { j = idx % BN; k = idx / BN; } else { k = idx % BK; j = idx / BK; }
            const int gj = j0 + j, gk = k0 + k;
            double v = 0.0;
            if (gj < N && gk < kend) v = TB ? B[(i64)gj + (i64)gk * ldb] : B[(i64)gk + (i64)gj * ldb];
            if ((mB && gk < gj) || (mBT && gk > gj)) v = 0.0;
            rb[l] = v;
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int l = 0; l < LA; l++) {
            const int idx = tid + l * NT;
            int i, k;
            if (!TA) { i = idx % BM; k = idx / BM; } else { k = idx % BK; i = idx / BK; }
            As[k][i] = ra[l];
        }
#pragma unroll
        for (int l = 0; l < LB; l++) {
            const int idx = tid + l * NT;
            int j, k;
            if (TB) { j = idx % BN; k = idx / BN; } else { k = idx % BK; j = idx / BK; }
            Bs[k][j] = rb[l];
        }
    };

    if (kbeg < kend) fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        __syncthreads();
        stage();
        __syncthreads();
        if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] = As[k4 + fk][wm + mi * 8 + fr];
#pragma unroll
            for (int ni = 0; ni < 4; ni++) b[ni] = Bs[k4 + fk][wn + ni * 8 + fr];
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma_8x8x4(acc[mi][ni], a[mi], b[ni]);
        }
    }

    const bool lower = (mode & GEMM_LOWER) != 0, mirror = (mode & GEMM_MIRROR) != 0;
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int gi = i0 + wm + mi * 8 + fr;
        if (gi >= M) continue;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int gj = j0 + wn + ni * 8 + 2 * fk + e;
                if (gj >= N) continue;
                if (lower && gj > gi) continue;
                double* c = C + (i64)gi + (i64)gj * ldc;
                double v = alpha * acc[mi][ni][e];
                if (beta != 0.0) v += beta * (*c);
                *c = v;
                if (mirror && gi != gj) C[(i64)gj + (i64)gi * ldc] = v;
            }
        }
    }
}

static void gemm_f64_mma_launch(cudaStream_t s, bool ta, bool tb, i64 M, i64 N, i64 K, double alpha, const double* A, i64 lda,
                                const double* B, i64 ldb, double beta, double* C, i64 ldc, int mode)
{
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)((N + 127) / 128));
    if (!ta && !tb) gemm_f64_mma_kernel<false, false><<<grid, 512, 0, s>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc, mode);
    else if (!ta && tb) gemm_f64_mma_kernel<false, true><<<grid, 512, 0, s>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc, mode);
    else if (ta && !tb) gemm_f64_mma_kernel<true, false><<<grid, 512, 0, s>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc, mode);
    else gemm_f64_mma_kernel<true, true><<<grid, 512, 0, s>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc, mode);
    KERNEL_CHECK();
}

template <class T, class Cfg>
static void gemm_launch(cudaStream_t s, bool ta, bool tb, i64 M, i64 N, i64 K, T alpha, const T* A, i64 lda,
                        const T* B, i64 ldb, T beta, T* C, i64 ldc, int mode)
{
    dim3 grid((unsigned)((M + Cfg::BM - 1) / Cfg::BM), (unsigned)((N + Cfg::BN - 1) / Cfg::BN));
    if (!ta && !tb) gemm_kernel<T, Cfg, false, false><<<grid, 256, 0, s>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc, mode);
    else if (!ta && tb) gemm_kernel<T, Cfg, false, true><<<grid, 256, 0, s>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc, mode);
    else if (ta && !tb) gemm_kernel<T, Cfg, true, false><<<grid, 256, 0, s>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc, mode);
    else gemm_kernel<T, Cfg, true, true><<<grid, 256, 0, s>>>((int)M, (int)N, (int)K, alpha, A, lda, B, ldb, beta, C, ldc, mode);
    KERNEL_CHECK();
}

template <class T>
void gemm(cudaStream_t s, bool ta, bool tb, i64 M, i64 N, i64 K, T alpha, const T* A, i64 lda,
          const T* B, i64 ldb, T beta, T* C, i64 ldc, int mode)
{
    if (M <= 0 || N <= 0) return;
    if (std::is_same<T, double>::value) {
        // float64: tensor-core tiles once the 128 x 128 grid fills the chip (B200ADMM_GEMM_F64=simt: CUDA cores)
        static int use_mma = -1;
        if (use_mma < 0) { const char* e = getenv("B200ADMM_GEMM_F64"); use_mma = (e && !strcmp(e, "simt")) ? 0 : 1; }
        const i64 tiles = ((M + 127) / 128) * ((N + 127) / 128);
        if (use_mma && K >= 16 && tiles >= (i64)sm_count()) {
            gemm_f64_mma_launch(s, ta, tb, M, N, K, (double)alpha, (const double*)A, lda, (const double*)B, ldb, (double)beta, (double*)C, ldc, mode);
            return;
        }
    }
    typedef typename GemmCfg<T>::Large L;
    typedef typename GemmCfg<T>::Small S;
    // narrow or small outputs (panels of the factorisation / triangular inverse): the large tile would
    // leave most SMs idle, so fall back to the 64 x 64 tile when the large-tile grid cannot fill the chip
    const i64 big_tiles = ((M + L::BM - 1) / L::BM) * ((N + L::BN - 1) / L::BN);
    if (big_tiles < 2 * (i64)sm_count()) gemm_launch<T, S>(s, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, mode);
    else gemm_launch<T, L>(s, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, mode);
}

template void gemm<float>(cudaStream_t, bool, bool, i64, i64, i64, float, const float*, i64, const float*, i64, float, float*, i64, int);
template void gemm<double>(cudaStream_t, bool, bool, i64, i64, i64, double, const double*, i64, const double*, i64, double, double*, i64, int);

}  // namespace b200
