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
__global__ void __launch_bounds__(DIAG_THREADS) chol_diag_kernel(T* __restrict__ A, i64 lda, int kb, int k0, T* __restrict__ Dinv, int* info, int winv_in_smem)
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
            chol_diag_kernel<T><<<1, DIAG_THREADS, smem, s>>>(Akk, lda, kb, (int)j0, Dinv + b * NB * NB, info_dev, winv_in_smem);
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

template <class T>
void chol_solve_vec(cudaStream_t s, const T* L, i64 p, i64 lda, T* b)
{
    chol_solve_vec_kernel<T><<<1, 1024, 0, s>>>(L, (int)p, lda, b);
    KERNEL_CHECK();
}
template void chol_solve_vec<float>(cudaStream_t, const float*, i64, i64, float*);
template void chol_solve_vec<double>(cudaStream_t, const double*, i64, i64, double*);

}  // namespace b200
