// synth.cu -- synthetic Gaussian design of the reference's benchmarks, generated in HBM.
//
// Mirrors the README recipe (/root/reference/README.md:195-201): X_ij ~ N(mean, sd^2) i.i.d.,
// beta* = nsig leading U(0,1) coefficients, y = X beta* + N(0, noise^2).  The generator is
// counter based (Philox4x32-10, key = seed, counter = (column, global row / 4)), so any row
// block of the same global matrix can be produced independently on any GPU -- that is what
// lets the row-sharded multi-GPU benchmark build its shards without moving data.
#include "common.cuh"
#include "kernels.h"

namespace b200 {

namespace {

struct U4 { unsigned x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1)
{
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const unsigned hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}
__device__ __forceinline__ float u01(unsigned x) { return ((float)x + 0.5f) * 2.3283064365386963e-10f; }
__device__ __forceinline__ void box_muller(unsigned a, unsigned b, float& n0, float& n1)
{
    const float u1 = fmaxf(u01(a), 1e-12f), u2 = u01(b);
    const float r = sqrtf(-2.f * logf(u1));
    float s, c;
    sincospif(2.f * u2, &s, &c);
    n0 = r * c; n1 = r * s;
}

// stream ids mixed into the counter's 4th word
constexpr unsigned STREAM_X = 0u, STREAM_NOISE = 0x5EEDu, STREAM_BETA = 0xBE7Au;

__global__ void __launch_bounds__(256) synth_x_kernel(float* __restrict__ X, i64 nrows, i64 p, i64 row0, unsigned k0, unsigned k1,
                                                      float mean_x, float sd_x)
{
    const i64 g_first = row0 / 4, g_last = (row0 + nrows + 3) / 4;      // global row groups touched
    const i64 ng = g_last - g_first;
    const i64 total = ng * p;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const i64 col = idx / ng, g = g_first + idx % ng;
        const U4 r = philox4x32_10((unsigned)col, (unsigned)(col >> 32), (unsigned)g, STREAM_X ^ (unsigned)(g >> 32), k0, k1);
        float v[4];
        box_muller(r.x, r.y, v[0], v[1]);
        box_muller(r.z, r.w, v[2], v[3]);
        const i64 lr = g * 4 - row0;                                    // local row of the group's first element
        float* dst = X + col * nrows;
        if (lr >= 0 && lr + 3 < nrows && (((uintptr_t)(dst + lr)) & 15) == 0) {
            float4 o = make_float4(mean_x + sd_x * v[0], mean_x + sd_x * v[1], mean_x + sd_x * v[2], mean_x + sd_x * v[3]);
            __stcs(reinterpret_cast<float4*>(dst + lr), o);
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (lr + e >= 0 && lr + e < nrows) dst[lr + e] = mean_x + sd_x * v[e];
        }
    }
}

__global__ void __launch_bounds__(256) synth_y_kernel(const float* __restrict__ X, float* __restrict__ y, i64 nrows, i64 p, i64 row0,
                                                      unsigned k0, unsigned k1, int nsig, float noise)
{
    __shared__ float beta[1024];
    for (int j = threadIdx.x; j < nsig; j += blockDim.x) {
        const U4 r = philox4x32_10((unsigned)j, 0u, 0u, STREAM_BETA, k0, k1);
        beta[j] = u01(r.x);
    }
    __syncthreads();
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const i64 gr = row0 + i;
    const U4 r = philox4x32_10((unsigned)(gr >> 1), (unsigned)(gr >> 33), 0u, STREAM_NOISE, k0, k1);
    float e0, e1;
    box_muller(r.x, r.y, e0, e1);
    float acc = 0.f;
    const int ns = (int)min((i64)nsig, p);
    for (int j = 0; j < ns; j++) acc = fmaf(X[i + (i64)j * nrows], beta[j], acc);
    y[i] = acc + noise * ((gr & 1) ? e1 : e0);
}

}  // namespace

void synth_design_f32(cudaStream_t s, float* X, float* y, i64 nrows, i64 p, i64 row0, uint64_t seed,
                      float mean_x, float sd_x, int nsig, float noise)
{
    if (nsig > 1024) throw ArgError("synth: nsig must be <= 1024");
    const unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
    const i64 total = ((nrows + 6) / 4 + 1) * p;
    const unsigned grid = (unsigned)std::min<i64>((total + 255) / 256, (i64)sm_count() * 32);
    synth_x_kernel<<<grid, 256, 0, s>>>(X, nrows, p, row0, k0, k1, mean_x, sd_x);
    KERNEL_CHECK();
    if (y) {
        synth_y_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, s>>>(X, y, nrows, p, row0, k0, k1, nsig, noise);
        KERNEL_CHECK();
    }
}

}  // namespace b200
