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

}  // namespace

// work layout: [nblk * NB * NB] inverse diagonal blocks, then [p * NB] panel scratch
template <class T> size_t chol_work(i64 p)
{
    const i64 nblk = (p + NB - 1) / NB;
    return (size_t)(nblk * NB * NB + p * NB);
}
template size_t chol_work<float>(i64);
template size_t chol_work<double>(i64);

template <class T>
void chol_lower(cudaStream_t s, T* A, i64 p, i64 lda, T* work, int* info_dev)
{
    const i64 nblk = (p + NB - 1) / NB;
    T* Dinv = work;
    T* panel = work + nblk * NB * NB;
    const int winv_in_smem = sizeof(T) == 4 ? 1 : 0;             // two 128 x 129 tiles: 132 KB in float, too large in double
    const size_t smem = sizeof(T) * NB * (NB + 1) * (winv_in_smem ? 2 : 1);
    static bool attr_done_f = false, attr_done_d = false;
    bool& done = std::is_same<T, float>::value ? attr_done_f : attr_done_d;
    if (!done) {
        CUDA_CHECK(cudaFuncSetAttribute(chol_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done = true;
    }
    CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int), s));
    for (i64 b = 0; b < nblk; b++) {
        const i64 k0 = b * NB;
        const int kb = (int)std::min<i64>(NB, p - k0);
        T* Akk = A + k0 + k0 * lda;
        chol_diag_kernel<T><<<1, DIAG_THREADS, smem, s>>>(Akk, lda, kb, (int)k0, Dinv + b * NB * NB, info_dev, winv_in_smem);
        KERNEL_CHECK();
        const i64 m = p - k0 - kb;
        if (m <= 0) break;
        T* A21 = A + (k0 + kb) + k0 * lda;
        // L21 = A21 * inv(L11)'
        gemm<T>(s, false, true, m, kb, kb, T(1), A21, lda, Dinv + b * NB * NB, NB, T(0), panel, m, 0);
        copy_block<T>(s, panel, m, A21, lda, m, kb);
        // A22 -= L21 L21'
        T* A22 = A + (k0 + kb) + (k0 + kb) * lda;
        gemm<T>(s, false, true, m, m, kb, T(-1), panel, m, panel, m, T(1), A22, lda, GEMM_LOWER);
    }
}
template void chol_lower<float>(cudaStream_t, float*, i64, i64, float*, int*);
template void chol_lower<double>(cudaStream_t, double*, i64, i64, double*, int*);

// W <- L^-1.  `work` is chol_lower's workspace (inverse diagonal blocks); tmp: p * NB entries.
template <class T>
void tri_inverse_lower(cudaStream_t s, const T* L, i64 p, i64 lda, const T* work, T* W, i64 ldw, T* tmp)
{
    const i64 nblk = (p + NB - 1) / NB;
    CUDA_CHECK(cudaMemsetAsync(W, 0, sizeof(T) * (size_t)ldw * (size_t)p, s));
    for (i64 b = nblk - 1; b >= 0; b--) {
        const i64 j0 = b * NB;
        const int kb = (int)std::min<i64>(NB, p - j0);
        const T* Dinv = work + b * NB * NB;
        copy_block<T>(s, Dinv, NB, W + j0 + j0 * ldw, ldw, kb, kb);
        const i64 m = p - j0 - kb;
        if (m <= 0) continue;
        // T = L21 * inv(L11)
        gemm<T>(s, false, false, m, kb, kb, T(1), L + (j0 + kb) + j0 * lda, lda, Dinv, NB, T(0), tmp, m, 0);
        // W21 = - W22 * T      (W22 lower triangular, already final)
        gemm<T>(s, false, false, m, kb, m, T(-1), W + (j0 + kb) + (j0 + kb) * ldw, ldw, tmp, m, T(0),
                W + (j0 + kb) + j0 * ldw, ldw, GEMM_A_LOWER_TRI);
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
