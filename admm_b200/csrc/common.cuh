// common.cuh -- shared helpers for the sm_100a kernels of libb200admm.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <stdexcept>
#include <utility>
#include <vector>

namespace b200 {

typedef long long i64;

// ------------------------------------------------------------------ error plumbing
struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct ArgError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct CodeError : std::runtime_error {      // carries one of the B200ADMM_E* codes
    int code;
    CodeError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line)
{
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
        throw CudaError(buf);
    }
}
#define CUDA_CHECK(x) ::b200::cuda_check((x), #x, __FILE__, __LINE__)
extern double g_last_gram_seconds;             // device time of the last tall solver's Gram kernel launches
extern double g_last_work[4];                  // b200admm_last_work: {algorithmic iteration bytes, regular steps, active-set steps, 0} of the last wide fit
extern unsigned long long g_launch_count;       // kernels launched by this library (capi.cu)
#define KERNEL_CHECK() do { ++::b200::g_launch_count; ::b200::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__); } while (0)

// ------------------------------------------------------------------ device buffers
// Large blocks (>= 64 MB) are recycled between calls through a small exact-size cache (capi.cu):
// cudaMalloc / cudaFree of the 40 GB working copy cost ~0.1 s per fit otherwise.
void* dev_alloc(size_t bytes);
void dev_free(void* p, size_t bytes);
void dev_cache_release();

template <class T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void alloc(size_t n_)
    {
        release();
        n = n_;
        if (n) p = (T*)dev_alloc(n * sizeof(T));
    }
    void release()
    {
        if (p) dev_free(p, n * sizeof(T));
        p = nullptr; n = 0;
    }
    void zero(cudaStream_t s) { if (n) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

struct EventTimer {                 // CUDA-event stopwatch on one stream
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    explicit EventTimer(cudaStream_t s_) : s(s_)
    {
        CUDA_CHECK(cudaEventCreate(&a));
        CUDA_CHECK(cudaEventCreate(&b));
    }
    ~EventTimer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    void start() { CUDA_CHECK(cudaEventRecord(a, s)); }
    double stop()   // seconds; synchronises the stream
    {
        CUDA_CHECK(cudaEventRecord(b, s));
        CUDA_CHECK(cudaEventSynchronize(b));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
        return ms * 1e-3;
    }
};

struct SpanTimer {                  // sums the device time of several bracketed spans; no synchronisation until total()
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
    cudaStream_t s;
    explicit SpanTimer(cudaStream_t s_) : s(s_) {}
    ~SpanTimer() { for (auto& e : spans) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); } }
    void begin()
    {
        cudaEvent_t a, b;
        CUDA_CHECK(cudaEventCreate(&a));
        CUDA_CHECK(cudaEventCreate(&b));
        spans.push_back({a, b});
        CUDA_CHECK(cudaEventRecord(a, s));
    }
    void end() { CUDA_CHECK(cudaEventRecord(spans.back().second, s)); }
    double total()   // seconds; every span must have completed (call after a stream synchronisation)
    {
        double sum = 0;
        for (auto& e : spans) {
            float ms = 0;
            CUDA_CHECK(cudaEventSynchronize(e.second));
            CUDA_CHECK(cudaEventElapsedTime(&ms, e.first, e.second));
            sum += ms * 1e-3;
        }
        return sum;
    }
};

int sm_count();                     // SMs of the current device (cached)

#ifdef __CUDACC__
// ------------------------------------------------------------------ device helpers
template <class T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; result valid in every thread.  `scratch` needs >= 33 entries of T.
template <class T> __device__ __forceinline__ T block_sum(T v, T* scratch)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();                 // protect scratch from the previous use
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        T t = lane < nw ? scratch[lane] : T(0);
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// streaming 128-bit load that does not allocate in L1 (matrix data read once per pass)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_stream_d2(const double2* p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                 : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
#endif

}  // namespace b200
