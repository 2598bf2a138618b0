// stdize.cu -- DataStd on the device (/root/reference/src/DataStd.h:89-155, non-AVX branch),
// plus the f64 -> f32 ingest conversion (/root/reference/src/Lasso.cpp:45-50).
//
// flag = standardize + 2 * intercept.  Per column (Scalar T = float for lasso/enet, double for LAD):
//   flag 1: scale = ||x - mean(x)|| / sqrt(n);            x *= 1/scale          (not centred)
//   flag 2: mean = mean(x);  x -= mean
//   flag 3: mean = mean(x);  x -= mean;  scale = ||x|| * (1/sqrt(n));  x *= 1/scale
// y is treated the same way except that it is *divided* by its scale.
//
// The column passes are split into three kernels (sum; centred sum of squares; apply) so that
// the row-sharded multi-GPU path can all-reduce the p-vectors between them; on one GPU they
// simply run back to back: 3 reads + 1 write of X, all coalesced 128-bit streams, one column
// per CTA at a time so each CTA walks contiguous memory.
#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

constexpr int STD_THREADS = 512;

template <class T> __device__ __forceinline__ T ldv(const T* p) { return *p; }

template <class T>
__global__ void __launch_bounds__(STD_THREADS) col_sum_kernel(const T* __restrict__ X, i64 n, i64 p, i64 ld, T* __restrict__ sums)
{
    __shared__ T scratch[33];
    for (i64 j = blockIdx.x; j < p; j += gridDim.x) {
        const T* c = X + j * ld;
        T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        i64 i = threadIdx.x;
        for (; i + 3 * STD_THREADS < n; i += 4 * STD_THREADS) {
            s0 += c[i]; s1 += c[i + STD_THREADS]; s2 += c[i + 2 * STD_THREADS]; s3 += c[i + 3 * STD_THREADS];
        }
        for (; i < n; i += STD_THREADS) s0 += c[i];
        T s = block_sum((s0 + s1) + (s2 + s3), scratch);
        if (threadIdx.x == 0) sums[j] = s;
    }
}

// sumsq[j] = sum_i (x_ij - mean_j)^2   (mean == nullptr -> 0); optionally stores the centred values
template <class T>
__global__ void __launch_bounds__(STD_THREADS) col_sumsq_kernel(const T* __restrict__ Xin, T* Xout, i64 n, i64 p, i64 ld,
                                                                const T* __restrict__ mean, T* __restrict__ sumsq)
{
    __shared__ T scratch[33];
    for (i64 j = blockIdx.x; j < p; j += gridDim.x) {
        const T* c = Xin + j * ld;
        T* o = Xout ? Xout + j * ld : nullptr;
        const T mu = mean ? mean[j] : T(0);
        T s0 = 0, s1 = 0;
        i64 i = threadIdx.x;
        for (; i + STD_THREADS < n; i += 2 * STD_THREADS) {
            const T a = c[i] - mu, b = c[i + STD_THREADS] - mu;
            if (o) { o[i] = a; o[i + STD_THREADS] = b; }
            s0 += a * a; s1 += b * b;
        }
        for (; i < n; i += STD_THREADS) { const T a = c[i] - mu; if (o) o[i] = a; s0 += a * a; }
        T s = block_sum(s0 + s1, scratch);
        if (threadIdx.x == 0) sumsq[j] = s;
    }
}

// out = (x - mean) * factor   or   (x - mean) / divisor;  null pointers skip the step
template <class T>
__global__ void __launch_bounds__(STD_THREADS) col_apply_kernel(const T* Xin, T* Xout, i64 n, i64 p, i64 ld, i64 ld_out,
                                                                const T* __restrict__ mean, const T* __restrict__ factor,
                                                                const T* __restrict__ divisor)
{
    for (i64 j = blockIdx.x; j < p; j += gridDim.x) {
        const T* c = Xin + j * ld;
        T* o = Xout + j * ld_out;
        const T mu = mean ? mean[j] : T(0);
        const T f = factor ? factor[j] : T(1);
        const T d = divisor ? divisor[j] : T(1);
        for (i64 i = threadIdx.x; i < n; i += STD_THREADS) {
            T v = c[i];
            if (mean) v = v - mu;
            if (factor) v = v * f;
            if (divisor) v = v / d;
            o[i] = v;
        }
    }
}

// tiny per-column arithmetic between the passes
template <class T>
__global__ void mean_from_sum_kernel(const T* sums, i64 p, i64 n_total, T* mean)
{
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < p) mean[j] = sums[j] / T(n_total);
}
// flag 3 / y: scale = sqrt(sumsq) * n_invsqrt;   flag 1: scale = sqrt(sumsq) / sqrt(n)
template <class T>
__global__ void scale_from_sumsq_kernel(const T* sumsq, i64 p, i64 n_total, int sd_form, T* scale, T* inv)
{
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    const T nrm = sqrt(sumsq[j]);
    T sc;
    if (sd_form) sc = nrm / sqrt(T(n_total));
    else {
        const T n_invsqrt = T(1.0 / (double)sqrt(T(n_total)));
        sc = nrm * n_invsqrt;
    }
    scale[j] = sc;
    if (inv) inv[j] = T(1.0 / (double)sc);
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, size_t count)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) out[i] = (float)in[i];
}

inline unsigned col_grid(i64 p) { return (unsigned)std::min<i64>(p, (i64)1 << 20); }   // one CTA per column, dynamic scheduling

}  // namespace

template <class T> void column_sums(cudaStream_t s, const T* X, i64 n, i64 p, i64 ld, T* sums)
{
    if (p <= 0) return;
    col_sum_kernel<T><<<col_grid(p), STD_THREADS, 0, s>>>(X, n, p, ld, sums);
    KERNEL_CHECK();
}
template <class T> void column_center_sumsq(cudaStream_t s, T* X, i64 n, i64 p, i64 ld, const T* mean, T* sumsq, bool center_in_place)
{
    if (p <= 0) return;
    col_sumsq_kernel<T><<<col_grid(p), STD_THREADS, 0, s>>>(X, center_in_place ? X : nullptr, n, p, ld, mean, sumsq);
    KERNEL_CHECK();
}
template <class T> void column_scale(cudaStream_t s, T* X, i64 n, i64 p, i64 ld, const T* inv_scale)
{
    if (p <= 0) return;
    col_apply_kernel<T><<<col_grid(p), STD_THREADS, 0, s>>>(X, X, n, p, ld, ld, nullptr, inv_scale, nullptr);
    KERNEL_CHECK();
}

template <class T> void mean_from_sums(cudaStream_t s, const T* sums, i64 p, i64 n_total, T* mean)
{
    mean_from_sum_kernel<T><<<(unsigned)((p + 255) / 256), 256, 0, s>>>(sums, p, n_total, mean);
    KERNEL_CHECK();
}
template <class T> void scale_from_sumsq(cudaStream_t s, const T* sumsq, i64 p, i64 n_total, bool sd_form, T* scale, T* inv)
{
    scale_from_sumsq_kernel<T><<<(unsigned)((p + 255) / 256), 256, 0, s>>>(sumsq, p, n_total, sd_form ? 1 : 0, scale, inv);
    KERNEL_CHECK();
}
template <class T> void column_apply(cudaStream_t s, const T* Xin, T* Xout, i64 n, i64 p, i64 ld, i64 ld_out, const T* mean, const T* factor, const T* divisor)
{
    if (p <= 0) return;
    col_apply_kernel<T><<<col_grid(p), STD_THREADS, 0, s>>>(Xin, Xout, n, p, ld, ld_out, mean, factor, divisor);
    KERNEL_CHECK();
}

// Single-device DataStd for the columns of X.  `tmp` : 2 * p scratch entries.
template <class T>
void standardize_columns(cudaStream_t s, const T* X_in, T* X_out, i64 n, i64 p, i64 ld, int flag, T* meanX, T* scaleX, T* tmp)
{
    T* sums = tmp;
    T* inv = tmp + p;
    switch (flag) {
    case 1:
        column_sums(s, X_in, n, p, ld, sums);
        mean_from_sums(s, sums, p, n, sums);                       // mean, used only for the deviation
        col_sumsq_kernel<T><<<col_grid(p), STD_THREADS, 0, s>>>(X_in, nullptr, n, p, ld, sums, scaleX); KERNEL_CHECK();
        scale_from_sumsq(s, scaleX, p, n, true, scaleX, inv);
        column_apply<T>(s, X_in, X_out, n, p, ld, ld, nullptr, inv, nullptr);
        break;
    case 2:
        column_sums(s, X_in, n, p, ld, sums);
        mean_from_sums(s, sums, p, n, meanX);
        column_apply<T>(s, X_in, X_out, n, p, ld, ld, meanX, nullptr, nullptr);
        break;
    case 3:
        column_sums(s, X_in, n, p, ld, sums);
        mean_from_sums(s, sums, p, n, meanX);
        col_sumsq_kernel<T><<<col_grid(p), STD_THREADS, 0, s>>>(X_in, nullptr, n, p, ld, meanX, scaleX); KERNEL_CHECK();
        scale_from_sumsq(s, scaleX, p, n, false, scaleX, inv);
        column_apply<T>(s, X_in, X_out, n, p, ld, ld, meanX, inv, nullptr);
        break;
    default:
        if (X_in != X_out) column_apply<T>(s, X_in, X_out, n, p, ld, ld, nullptr, nullptr, nullptr);
        break;
    }
}

// y in place; out2 = {meanY, scaleY} (device).  `tmp`: 2 scratch entries.
template <class T> void standardize_y(cudaStream_t s, T* y, i64 n, int flag, T* out2, T* tmp)
{
    T* meanY = out2;
    T* scaleY = out2 + 1;
    switch (flag) {
    case 1:
        column_sums(s, y, n, 1, n, tmp);
        mean_from_sums(s, tmp, 1, n, tmp);
        col_sumsq_kernel<T><<<1, STD_THREADS, 0, s>>>(y, nullptr, n, 1, n, tmp, scaleY); KERNEL_CHECK();
        scale_from_sumsq<T>(s, scaleY, 1, n, true, scaleY, nullptr);
        column_apply<T>(s, y, y, n, 1, n, n, nullptr, nullptr, scaleY);
        break;
    case 2:
    case 3:
        column_sums(s, y, n, 1, n, tmp);
        mean_from_sums(s, tmp, 1, n, meanY);
        col_sumsq_kernel<T><<<1, STD_THREADS, 0, s>>>(y, nullptr, n, 1, n, meanY, scaleY); KERNEL_CHECK();
        scale_from_sumsq<T>(s, scaleY, 1, n, false, scaleY, nullptr);
        column_apply<T>(s, y, y, n, 1, n, n, meanY, nullptr, scaleY);
        break;
    default:
        break;
    }
}

void convert_f64_to_f32(cudaStream_t s, const double* in, float* out, size_t count)
{
    if (!count) return;
    const unsigned grid = (unsigned)std::min<size_t>((count + 1023) / 1024, (size_t)sm_count() * 16);
    f64_to_f32_kernel<<<grid, 256, 0, s>>>(in, out, count);
    KERNEL_CHECK();
}

#define INST(T)                                                                                              \
    template void column_sums<T>(cudaStream_t, const T*, i64, i64, i64, T*);                                 \
    template void column_center_sumsq<T>(cudaStream_t, T*, i64, i64, i64, const T*, T*, bool);               \
    template void column_scale<T>(cudaStream_t, T*, i64, i64, i64, const T*);                                \
    template void mean_from_sums<T>(cudaStream_t, const T*, i64, i64, T*);                                   \
    template void scale_from_sumsq<T>(cudaStream_t, const T*, i64, i64, bool, T*, T*);                       \
    template void column_apply<T>(cudaStream_t, const T*, T*, i64, i64, i64, i64, const T*, const T*, const T*);  \
    template void standardize_columns<T>(cudaStream_t, const T*, T*, i64, i64, i64, int, T*, T*, T*);        \
    template void standardize_y<T>(cudaStream_t, T*, i64, int, T*, T*);
INST(float)
INST(double)

}  // namespace b200
