// gemv.cu -- bandwidth-bound matrix-vector products (f32 / f64), column-major, deterministic.
//
//   gemv_t : out = A' v   one dot product per (contiguous) column.  Used for X'y
//            (/root/reference/src/ADMMLassoTall.h:172), the symmetric products of the coarse
//            Lanczos (src/Spectra/MatOp/DenseSymMatProd.h:64-69 -- the Gram matrix is stored
//            full, so A'v == Av), LAD's X'v (src/ADMMLAD.h:75) and BP's M'(.) (src/ADMMBP.h:66).
//   gemv_n : out = A v    row-parallel with a fixed-order reduction over column chunks.  Used
//            for LAD's X t (src/ADMMLAD.h:77) and BP's M v (src/ADMMBP.h:65).
//
// HBM-bound: every matrix element is read exactly once with 128-bit streaming loads that
// bypass L1; the vector operand stays in L1/L2.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace b200 {

template <class T> struct Vec;
template <> struct Vec<float>  { typedef float4 type;  static constexpr int N = 4; };
template <> struct Vec<double> { typedef double2 type; static constexpr int N = 2; };

__device__ __forceinline__ float4 ld_stream(const float4* p) { return ld_stream_f4(p); }
__device__ __forceinline__ double2 ld_stream(const double2* p) { return ld_stream_d2(p); }
__device__ __forceinline__ float dot_acc(float4 a, float4 b, float s)
{
    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    return s;
}
__device__ __forceinline__ double dot_acc(double2 a, double2 b, double s)
{
    s = fma(a.x, b.x, s); s = fma(a.y, b.y, s);
    return s;
}

// partial dot product of a[0..m) and v[0..m) over the lanes [lane, lane + nl, ...)
template <class T>
__device__ __forceinline__ T strided_dot(const T* __restrict__ a, const T* __restrict__ v, i64 m, int lane, int nl)
{
    typedef typename Vec<T>::type VT;
    constexpr int VN = Vec<T>::N;
    T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    const bool aligned = ((((uintptr_t)a) | ((uintptr_t)v)) & 15) == 0;
    if (aligned) {
        const i64 mv = m / VN;
        const VT* av = reinterpret_cast<const VT*>(a);
        const VT* vv = reinterpret_cast<const VT*>(v);
        i64 i = lane;
        for (; i + 3 * (i64)nl < mv; i += 4 * (i64)nl) {
            VT a0 = ld_stream(av + i), a1 = ld_stream(av + i + nl), a2 = ld_stream(av + i + 2 * (i64)nl), a3 = ld_stream(av + i + 3 * (i64)nl);
            VT b0 = vv[i], b1 = vv[i + nl], b2 = vv[i + 2 * (i64)nl], b3 = vv[i + 3 * (i64)nl];
            s0 = dot_acc(a0, b0, s0); s1 = dot_acc(a1, b1, s1); s2 = dot_acc(a2, b2, s2); s3 = dot_acc(a3, b3, s3);
        }
        for (; i < mv; i += nl) s0 = dot_acc(ld_stream(av + i), vv[i], s0);
        for (i64 r = mv * VN + lane; r < m; r += nl) s1 += a[r] * v[r];
    } else {
        i64 i = lane;
        for (; i + 3 * (i64)nl < m; i += 4 * (i64)nl) {
            s0 += a[i] * v[i]; s1 += a[i + nl] * v[i + nl];
            s2 += a[i + 2 * (i64)nl] * v[i + 2 * (i64)nl]; s3 += a[i + 3 * (i64)nl] * v[i + 3 * (i64)nl];
        }
        for (; i < m; i += nl) s0 += a[i] * v[i];
    }
    return (s0 + s1) + (s2 + s3);
}

// one warp per column
template <class T>
__global__ void __launch_bounds__(256) gemv_t_warp_kernel(const T* __restrict__ A, i64 m, i64 ncol, i64 lda,
                                                          const T* __restrict__ v, T* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const i64 nwarps = (i64)gridDim.x * (blockDim.x >> 5);
    for (i64 j = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < ncol; j += nwarps) {
        T s = strided_dot(A + j * lda, v, m, lane, 32);
        s = warp_sum(s);
        if (lane == 0) out[j] = s;
    }
}

// one block per column (long columns)
template <class T>
__global__ void __launch_bounds__(256) gemv_t_block_kernel(const T* __restrict__ A, i64 m, i64 ncol, i64 lda,
                                                           const T* __restrict__ v, T* __restrict__ out)
{
    __shared__ T scratch[33];
    for (i64 j = blockIdx.x; j < ncol; j += gridDim.x) {
        T s = strided_dot(A + j * lda, v, m, (int)threadIdx.x, (int)blockDim.x);
        s = block_sum(s, scratch);
        if (threadIdx.x == 0) out[j] = s;
    }
}

// long columns cut into nseg row segments: one CTA per (column, segment), partial sums added in segment order
template <class T>
__global__ void __launch_bounds__(256) gemv_t_seg_kernel(const T* __restrict__ A, i64 m, i64 lda, const T* __restrict__ v,
                                                         int nseg, i64 seg_len, T* __restrict__ partial)
{
    __shared__ T scratch[33];
    const i64 j = blockIdx.x / nseg;
    const int seg = (int)(blockIdx.x % nseg);
    const i64 r0 = seg * seg_len, r1 = min(m, r0 + seg_len);
    T s = r1 > r0 ? strided_dot(A + j * lda + r0, v + r0, r1 - r0, (int)threadIdx.x, (int)blockDim.x) : T(0);
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
template <class T>
__global__ void __launch_bounds__(256) gemv_t_seg_finish_kernel(const T* __restrict__ partial, i64 ncol, int nseg, T* __restrict__ out)
{
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncol) return;
    T s = partial[j * nseg];
    for (int q = 1; q < nseg; q++) s += partial[j * nseg + q];
    out[j] = s;
}

// the listed columns only: out[list[k]] = A(:, list[k])' v, one warp per entry, with the SAME summation order as
// gemv_t_warp_kernel (bit-identical values; the wide solver's screened regular step relies on it)
template <class T>
__global__ void __launch_bounds__(256) gemv_t_list_kernel(const T* __restrict__ A, i64 m, i64 lda, const T* __restrict__ v,
                                                          const int* __restrict__ list, int count, T* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = (int)(gridDim.x * (blockDim.x >> 5));
    for (int k = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)); k < count; k += nwarps) {
        const i64 j = list[k];
        T s = strided_dot(A + j * lda, v, m, lane, 32);
        s = warp_sum(s);
        if (lane == 0) out[j] = s;
    }
}
bool gemv_t_uses_warp_kernel(i64 m) { return m < 32768; }
template <class T>
void gemv_t_list(cudaStream_t s, const T* A, i64 m, i64 lda, const T* v, const int* list, int count, T* out)
{
    if (count <= 0) return;
    const i64 blocks = std::min<i64>(((i64)count + 7) / 8, (i64)sm_count() * 8);
    gemv_t_list_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(A, m, lda, v, list, count, out);
    KERNEL_CHECK();
}
template void gemv_t_list<float>(cudaStream_t, const float*, i64, i64, const float*, const int*, int, float*);

template <class T>
void gemv_t(cudaStream_t s, const T* A, i64 m, i64 ncol, i64 lda, const T* v, T* out)
{
    if (ncol <= 0) return;
    const int sms = sm_count();
    if (m >= 32768) {
        // One CTA per column leaves the last wave of resident CTAs mostly empty (5000 columns over 8 x 148 slots:
        // 4.2 waves cost 5; LAD's X'v ran at 67 % of the HBM rate).  Columns are therefore cut into row segments
        // until there are ~16 waves of CTAs; the segment sums are added in segment order (deterministic).
        const i64 slots = (i64)sms * 8;
        // (at least two segments: with one CTA per whole column the 8e4 x 8e4 product of the consensus solver ran at 5.39 TB/s,
        // cut in two at 6.77 TB/s -- profiles/r2x_gemv_nseg.log)
        const int nseg = (int)std::max<i64>(2, std::min<i64>(8, (16 * slots + ncol - 1) / ncol));
        if (ncol * nseg < ((i64)1 << 30)) {
            const i64 seg_len = (((m + nseg - 1) / nseg) + 3) & ~(i64)3;
            DevBuf<T> partial((size_t)ncol * nseg);
            gemv_t_seg_kernel<T><<<(unsigned)(ncol * nseg), 256, 0, s>>>(A, m, lda, v, nseg, seg_len, partial.p);
            KERNEL_CHECK();
            gemv_t_seg_finish_kernel<T><<<(unsigned)((ncol + 255) / 256), 256, 0, s>>>(partial.p, ncol, nseg, out);
            KERNEL_CHECK();
            return;      // (the block cache keeps `partial` alive for the stream-ordered kernels above)
        }
        const i64 grid = std::min<i64>(ncol, (i64)1 << 20);
        gemv_t_block_kernel<T><<<(unsigned)grid, 256, 0, s>>>(A, m, ncol, lda, v, out);
    } else {
        const i64 blocks = std::min<i64>((ncol + 7) / 8, (i64)sms * 8);
        gemv_t_warp_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(A, m, ncol, lda, v, out);
    }
    KERNEL_CHECK();
}
template void gemv_t<float>(cudaStream_t, const float*, i64, i64, i64, const float*, float*);
template void gemv_t<double>(cudaStream_t, const double*, i64, i64, i64, const double*, double*);

// ---------------------------------------------------------------------------------------
// out = A v, rows across threads, columns split into gridDim.y chunks
// ---------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) gemv_n_kernel(const T* __restrict__ A, i64 m, i64 ncol, i64 lda,
                                                     const T* __restrict__ v, T* __restrict__ part, i64 cols_per_chunk)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 j0 = (i64)blockIdx.y * cols_per_chunk;
    const i64 j1 = min(ncol, j0 + cols_per_chunk);
    __shared__ T vs[256];
    T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (i64 jb = j0; jb < j1; jb += 256) {
        const int cnt = (int)min((i64)256, j1 - jb);
        __syncthreads();
        if ((int)threadIdx.x < cnt) vs[threadIdx.x] = v[jb + threadIdx.x];
        __syncthreads();
        if (i < m) {
            const T* a = A + i + jb * lda;
            int j = 0;
            for (; j + 8 <= cnt; j += 8) {          // eight independent column loads in flight per thread
                T a0 = a[(i64)j * lda], a1 = a[(i64)(j + 1) * lda], a2 = a[(i64)(j + 2) * lda], a3 = a[(i64)(j + 3) * lda];
                T a4 = a[(i64)(j + 4) * lda], a5 = a[(i64)(j + 5) * lda], a6 = a[(i64)(j + 6) * lda], a7 = a[(i64)(j + 7) * lda];
                s0 += a0 * vs[j]; s1 += a1 * vs[j + 1]; s2 += a2 * vs[j + 2]; s3 += a3 * vs[j + 3];
                s0 += a4 * vs[j + 4]; s1 += a5 * vs[j + 5]; s2 += a6 * vs[j + 6]; s3 += a7 * vs[j + 7];
            }
            for (; j + 4 <= cnt; j += 4) {
                const T a0 = a[(i64)j * lda], a1 = a[(i64)(j + 1) * lda], a2 = a[(i64)(j + 2) * lda], a3 = a[(i64)(j + 3) * lda];
                s0 += a0 * vs[j]; s1 += a1 * vs[j + 1]; s2 += a2 * vs[j + 2]; s3 += a3 * vs[j + 3];
            }
            for (; j < cnt; j++) s0 += a[(i64)j * lda] * vs[j];
        }
    }
    if (i < m) part[(i64)blockIdx.y * m + i] = (s0 + s1) + (s2 + s3);
}
template <class T>
__global__ void gemv_n_reduce_kernel(const T* __restrict__ part, i64 m, int chunks, T* __restrict__ out)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    T s = 0;
    for (int c = 0; c < chunks; c++) s += part[(i64)c * m + i];
    out[i] = s;
}

static int gemv_n_chunks(i64 m, i64 ncol)
{
    const i64 bx = (m + 255) / 256;
    const i64 want = (i64)sm_count() * 6;
    i64 chunks = (want + bx - 1) / bx;
    chunks = std::max<i64>(1, std::min<i64>(chunks, (ncol + 63) / 64));
    return (int)std::min<i64>(chunks, 4096);
}
size_t gemv_n_work(i64 m, i64 ncol) { return (size_t)gemv_n_chunks(m, ncol) * (size_t)m; }

template <class T>
void gemv_n(cudaStream_t s, const T* A, i64 m, i64 ncol, i64 lda, const T* v, T* out, T* work)
{
    if (m <= 0) return;
    const int chunks = gemv_n_chunks(m, ncol);
    const i64 cpc = (ncol + chunks - 1) / chunks;
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)chunks);
    if (chunks == 1) {
        gemv_n_kernel<T><<<grid, 256, 0, s>>>(A, m, ncol, lda, v, out, cpc);
    } else {
        gemv_n_kernel<T><<<grid, 256, 0, s>>>(A, m, ncol, lda, v, work, cpc);
        gemv_n_reduce_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, s>>>(work, m, chunks, out);
    }
    KERNEL_CHECK();
}
template void gemv_n<float>(cudaStream_t, const float*, i64, i64, i64, const float*, float*, float*);
template void gemv_n<double>(cudaStream_t, const double*, i64, i64, i64, const double*, double*, double*);

}  // namespace b200
