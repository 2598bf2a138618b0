// capi.cu -- the C ABI of libb200admm.so (include/b200admm.h) and the host-side path drivers.
//
// Host logic restated from the reference's native entry points:
//   admm_lasso    /root/reference/src/Lasso.cpp:32-138     -> b200admm_lasso
//   admm_enet     /root/reference/src/Enet.cpp:31-138      -> b200admm_enet
//   DataStd::recover  /root/reference/src/DataStd.h:183-207 -> recover_sparse()
//   write_beta_matrix /root/reference/src/Lasso.cpp:22-30  -> CSC assembly in finish_path()
// All numerical work runs in the CUDA kernels of this directory; there is no CPU fallback.
#include "../../include/b200admm.h"
#include "common.cuh"
#include "kernels.h"
#include "coarse_eig.hpp"
#include "solvers.h"
#include "comm.h"

#include <chrono>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

namespace b200 {

static thread_local std::string g_last_error;
unsigned long long g_launch_count = 0;
double g_last_gram_seconds = 0.0;
double g_last_work[4] = {0, 0, 0, 0};

int sm_count()
{
    static int cached = 0;
    if (!cached) {
        int dev = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        CUDA_CHECK(cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev));
    }
    return cached;
}

Context& ctx()
{
    static Context c;
    if (!c.ready) {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw CodeError(B200ADMM_ENODEVICE, std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
        int dev = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
        if (prop.major < 10)
            throw CodeError(B200ADMM_ENODEVICE, std::string("device '") + prop.name + "' is not a Blackwell (sm_100a) GPU");
        int coop = 0;
        CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        if (!coop) throw CodeError(B200ADMM_ENODEVICE, "device does not support cooperative launch");
        CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        c.device = dev;
        c.ready = true;
    }
    return c;
}

// ---- exact-size recycling of device blocks ---------------------------------------------------
// Every solver call allocates the same set of sizes; cudaMalloc / cudaFree are driver calls that take
// global locks (and cudaFree synchronises the device), ~0.1 s per fit for the 40 GB working copies and a
// source of multi-millisecond stalls on a busy host even for small blocks.  Freed blocks are therefore
// kept and handed out again on an exact size match.  All work is ordered on the library's one stream
// (the copy stream is joined before its buffers are released), so reuse without a synchronisation is safe.
namespace {
struct CachedBlock { void* p; size_t bytes; };
std::vector<CachedBlock>& block_cache() { static std::vector<CachedBlock> c; return c; }
std::mutex& cache_mutex() { static std::mutex m; return m; }
constexpr size_t LARGE_BYTES = (size_t)64 << 20;
constexpr size_t MAX_LARGE = 16, MAX_SMALL = 256;
bool cache_enabled()
{
    static int on = -1;
    if (on < 0) { const char* e = getenv("B200ADMM_CACHE"); on = (e && !strcmp(e, "0")) ? 0 : 1; }
    return on == 1;
}
}  // namespace

void dev_cache_release()
{
    std::lock_guard<std::mutex> g(cache_mutex());
    for (auto& b : block_cache()) cudaFree(b.p);
    block_cache().clear();
}

void* dev_alloc(size_t bytes)
{
    if (cache_enabled()) {
        std::lock_guard<std::mutex> g(cache_mutex());
        auto& c = block_cache();
        for (size_t i = c.size(); i-- > 0;)
            if (c[i].bytes == bytes) { void* p = c[i].p; c.erase(c.begin() + i); return p; }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {          // make room: drop everything we are holding on to, retry once
        cudaGetLastError();
        dev_cache_release();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw CodeError(B200ADMM_ENOMEM, std::string("cudaMalloc of ") + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
    }
    return p;
}

void dev_free(void* p, size_t bytes)
{
    if (!p) return;
    if (cache_enabled()) {
        std::lock_guard<std::mutex> g(cache_mutex());
        auto& c = block_cache();
        size_t nlarge = 0;
        for (auto& b : c) nlarge += b.bytes >= LARGE_BYTES;
        const bool large = bytes >= LARGE_BYTES;
        if (large ? nlarge < MAX_LARGE : c.size() - nlarge < MAX_SMALL) { c.push_back({p, bytes}); return; }
    }
    cudaFree(p);
}

cudaStream_t copy_stream()
{
    Context& c = ctx();
    if (!c.copy) CUDA_CHECK(cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking));
    return c.copy;
}

TraceRequest& trace_request()
{
    static TraceRequest t;
    return t;
}

CaptureRequest& capture_request()
{
    static CaptureRequest c;
    return c;
}

double wall_now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
void make_lambda_grid(double lmax, double ratio, int nl, std::vector<double>& out)
{
    // lambda.setLinSpaced(nlambda, log(lmax), log(lmin)).exp()   (Lasso.cpp:86-88), Eigen 3.3 semantics:
    // one value -> `high` (so nlambda = 1 fits at lmin_ratio * lmax, not at lmax); otherwise the end nearer
    // zero is reached by stepping from the other end (low + i step, last = high; or first = low,
    // high - (n - 1 - i) step when |high| < |low|).
    out.resize(nl);
    const double lo = std::log(lmax), hi = std::log(ratio * lmax);
    if (nl == 1) { out[0] = std::exp(hi); return; }
    const double step = (hi - lo) / (nl - 1);
    const bool flip = std::fabs(hi) < std::fabs(lo);
    for (int i = 0; i < nl; i++) {
        double v;
        if (flip) v = (i == 0) ? lo : hi - (double)(nl - 1 - i) * step;
        else v = (i == nl - 1) ? hi : lo + (double)i * step;
        out[i] = std::exp(v);
    }
}

void free_path(b200admm_path* out)
{
    if (!out) return;
    free(out->lambda); free(out->niter); free(out->colptr); free(out->rowidx); free(out->val);
    out->lambda = nullptr; out->niter = nullptr; out->colptr = nullptr; out->rowidx = nullptr; out->val = nullptr;
}

// coefficient columns (standardised scale, zeros = absent) -> dgCMatrix pieces.
// with_intercept: row 0 holds beta0 (always stored), coefficient j goes to row j + 1.
template <class T>
void assemble_csc(const std::vector<std::vector<T>>& cols, const std::vector<T>& beta0, bool with_intercept,
                  i64 p, b200admm_path* out)
{
    const int nl = (int)cols.size();
    size_t nnz = 0;
    for (int k = 0; k < nl; k++) {
        if (with_intercept) nnz++;
        for (i64 j = 0; j < p; j++) if (cols[k][j] != T(0)) nnz++;
    }
    out->colptr = (int64_t*)malloc(sizeof(int64_t) * (nl + 1));
    out->rowidx = (int*)malloc(sizeof(int) * std::max<size_t>(nnz, 1));
    out->val = (double*)malloc(sizeof(double) * std::max<size_t>(nnz, 1));
    if (!out->colptr || !out->rowidx || !out->val) throw CodeError(B200ADMM_ENOMEM, "out of host memory");
    size_t pos = 0;
    for (int k = 0; k < nl; k++) {
        out->colptr[k] = (int64_t)pos;
        if (with_intercept) { out->rowidx[pos] = 0; out->val[pos] = (double)beta0[k]; pos++; }
        for (i64 j = 0; j < p; j++)
            if (cols[k][j] != T(0)) { out->rowidx[pos] = (int)(j + (with_intercept ? 1 : 0)); out->val[pos] = (double)cols[k][j]; pos++; }
    }
    out->colptr[nl] = (int64_t)pos;
    out->nrow = p + (with_intercept ? 1 : 0);
}
template void assemble_csc<float>(const std::vector<std::vector<float>>&, const std::vector<float>&, bool, i64, b200admm_path*);
template void assemble_csc<double>(const std::vector<std::vector<double>>&, const std::vector<double>&, bool, i64, b200admm_path*);

// DataStd::recover for a sparse coefficient vector (DataStd.h:183-207): only stored (non-zero)
// entries are touched, the inner product runs over them in index order, all in Scalar.
template <class T>
T recover_sparse(int flag, std::vector<T>& coef, const std::vector<T>& meanX, const std::vector<T>& scaleX, T meanY, T scaleY)
{
    const size_t p = coef.size();
    T beta0 = 0;
    switch (flag) {
    case 1:
        for (size_t j = 0; j < p; j++) if (coef[j] != T(0)) { coef[j] /= scaleX[j]; coef[j] *= scaleY; }
        break;
    case 2: {
        T s = 0;
        for (size_t j = 0; j < p; j++) if (coef[j] != T(0)) { coef[j] *= scaleY; s += coef[j] * meanX[j]; }
        beta0 = meanY - s;
    } break;
    case 3: {
        T s = 0;
        for (size_t j = 0; j < p; j++) if (coef[j] != T(0)) { coef[j] /= scaleX[j]; coef[j] *= scaleY; s += coef[j] * meanX[j]; }
        beta0 = meanY - s;
    } break;
    default: break;
    }
    return beta0;
}
template float recover_sparse<float>(int, std::vector<float>&, const std::vector<float>&, const std::vector<float>&, float, float);
template double recover_sparse<double>(int, std::vector<double>&, const std::vector<double>&, const std::vector<double>&, double, double);

// ---------------------------------------------------------------------------------------------
// ingest: caller's matrix -> library-owned float32 (or float64) device copy
// ---------------------------------------------------------------------------------------------
void ingest_f32(cudaStream_t s, const void* src, int dtype, size_t count, float* dst)
{
    if (!count) return;
    switch (dtype) {
    case B200ADMM_F32_DEVICE:
        CUDA_CHECK(cudaMemcpyAsync(dst, src, count * sizeof(float), cudaMemcpyDeviceToDevice, s));
        break;
    case B200ADMM_F32_HOST: {
        // chunked so that pageable sources pipeline through the driver's staging buffers
        const size_t chunk = (size_t)256 << 20;   // floats
        for (size_t o = 0; o < count; o += chunk) {
            const size_t c = std::min(chunk, count - o);
            CUDA_CHECK(cudaMemcpyAsync(dst + o, (const float*)src + o, c * sizeof(float), cudaMemcpyHostToDevice, s));
        }
    } break;
    case B200ADMM_F64_HOST: {
        // R's doubles: copy in slabs, narrow on the device (Lasso.cpp:45-50 does this on the CPU)
        const size_t chunk = (size_t)64 << 20;    // doubles per slab (512 MB)
        DevBuf<double> slab[2];
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; i++) { slab[i].alloc(std::min(chunk, count)); CUDA_CHECK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)); }
        int b = 0;
        for (size_t o = 0; o < count; o += chunk, b ^= 1) {
            const size_t c = std::min(chunk, count - o);
            CUDA_CHECK(cudaEventSynchronize(ev[b]));
            CUDA_CHECK(cudaMemcpyAsync(slab[b].p, (const double*)src + o, c * sizeof(double), cudaMemcpyHostToDevice, s));
            convert_f64_to_f32(s, slab[b].p, dst + o, c);
            CUDA_CHECK(cudaEventRecord(ev[b], s));
        }
        CUDA_CHECK(cudaStreamSynchronize(s));
        for (int i = 0; i < 2; i++) cudaEventDestroy(ev[i]);
    } break;
    case B200ADMM_F64_DEVICE:
        convert_f64_to_f32(s, (const double*)src, dst, count);
        break;
    default:
        throw ArgError("unknown dtype");
    }
}

void ingest_f64(cudaStream_t s, const void* src, int dtype, size_t count, double* dst)
{
    if (!count) return;
    switch (dtype) {
    case B200ADMM_F64_DEVICE:
        CUDA_CHECK(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToDevice, s));
        break;
    case B200ADMM_F64_HOST: {
        const size_t chunk = (size_t)128 << 20;
        for (size_t o = 0; o < count; o += chunk) {
            const size_t c = std::min(chunk, count - o);
            CUDA_CHECK(cudaMemcpyAsync(dst + o, (const double*)src + o, c * sizeof(double), cudaMemcpyHostToDevice, s));
        }
    } break;
    default:
        throw ArgError("this solver computes in float64: pass B200ADMM_F64_HOST or B200ADMM_F64_DEVICE data");
    }
}

__global__ void add_diag_kernel(float* A, i64 ld, i64 p, float v)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p) A[i + i * ld] += v;
}
void add_to_diagonal(cudaStream_t s, float* A, i64 ld, i64 p, float v)
{
    add_diag_kernel<<<(unsigned)((p + 255) / 256), 256, 0, s>>>(A, ld, p, v);
    KERNEL_CHECK();
}

// coarse lambda_max of the symmetric matrix S (full storage, n x n, lds) -- <= 13 device products
float coarse_eig_device(cudaStream_t s, const float* S, i64 n, i64 lds, int* nmatvec)
{
    if (n < 3) throw CodeError(B200ADMM_ELANCZOS, "coarse eigenvalue estimate needs at least 3 variables (Spectra: 1 <= nev < ncv <= n)");
    DevBuf<float> dv(n), dw(n);
    eig::LanczosInfo info;
    auto op = [&](const float* v, float* w) {
        CUDA_CHECK(cudaMemcpyAsync(dv.p, v, n * sizeof(float), cudaMemcpyHostToDevice, s));
        gemv_t<float>(s, S, n, n, lds, dv.p, dw.p);
        CUDA_CHECK(cudaMemcpyAsync(w, dw.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
    };
    const float ev = eig::coarse_largest_eigenvalue<float>(op, n, &info);
    if (nmatvec) *nmatvec = info.nmatvec;
    return ev;
}

// a <- a^-1 for the SPD matrix in `a` (p x p, ld), full symmetric result.  W: p x ld scratch,
// left holding L^-1 (lower triangular).  keep_factor: copy of the Cholesky factor L (p x ld) or null.
template <class T>
void spd_inverse(cudaStream_t s, T* a, i64 p, i64 ld, T* W, int* info_host, T* keep_factor)
{
    // The blocked factorisation and inverse are ~7 short launches per 128 columns (550 at p = 1e4), each
    // far shorter than a host hiccup: they are recorded into a CUDA graph and replayed as ONE launch, so
    // the device never waits for the host between them.  The instantiated graph (with its workspace) is
    // kept and replayed as long as the next call brings the same matrices (the block cache hands the same
    // pointers out again), which also removes the capture cost from every call but the first.  A
    // non-positive pivot does not stop the kernels (the pivot is replaced, `info` remembers the column);
    // it is reported after the graph has run.
    struct FactorGraph {
        cudaGraphExec_t exec = nullptr;
        const void *a = nullptr, *W = nullptr, *keep = nullptr;
        i64 p = -1, ld = -1;
        unsigned long long kernels = 0;       // kernel nodes in the graph (added to the launch count on every replay)
        DevBuf<T> work, tmp, zt;
        DevBuf<int> info;
        bool tensor = false;
    };
    static FactorGraph& fg = *new FactorGraph;      // lives until process exit (its buffers go back to the driver with the context)
    const char* genv = getenv("B200ADMM_GRAPH");
    const bool use_graph = !(genv && !strcmp(genv, "0"));
    const bool hit = use_graph && fg.exec && fg.a == a && fg.W == W && fg.keep == keep_factor && fg.p == p && fg.ld == ld;
    if (!hit) {
        if (fg.exec) { cudaGraphExecDestroy(fg.exec); fg.exec = nullptr; }
        // float32, p >= 512, no factor to hand back: everything on the tensor cores (chol.cu: spd_inverse_tc)
        const bool tensor = std::is_same<T, float>::value && !keep_factor && spd_inverse_tc_usable(p, ld);
        if (fg.p != p || fg.ld != ld || fg.tensor != tensor) {
            fg.work.release(); fg.tmp.release(); fg.zt.release();
            if (tensor) { fg.work.alloc(spd_tc_work_floats(p)); fg.zt.alloc((size_t)p * (size_t)ld); }
            else { fg.work.alloc(chol_work<T>(p)); fg.tmp.alloc(tri_inverse_tmp(p)); }
        }
        fg.tensor = tensor;
        if (!fg.info.p) fg.info.alloc(1);
        fg.a = a; fg.W = W; fg.keep = keep_factor; fg.p = p; fg.ld = ld;
        cudaGraph_t graph = nullptr;
        const unsigned long long count0 = g_launch_count;
        if (use_graph) CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
        try {
            if (tensor) {
                spd_inverse_tc(s, reinterpret_cast<float*>(a), p, ld, reinterpret_cast<float*>(W), reinterpret_cast<float*>(fg.zt.p),
                               reinterpret_cast<float*>(fg.work.p), fg.info.p);
            } else {
                chol_lower<T>(s, a, p, ld, fg.work.p, fg.info.p);
                if (keep_factor) CUDA_CHECK(cudaMemcpyAsync(keep_factor, a, sizeof(T) * (size_t)ld * (size_t)p, cudaMemcpyDeviceToDevice, s));
                tri_inverse_lower<T>(s, a, p, ld, fg.work.p, W, ld, fg.tmp.p);
                gram_of_lower<T>(s, W, p, ld, a, ld);
            }
        } catch (...) {
            if (use_graph) { cudaStreamEndCapture(s, &graph); if (graph) cudaGraphDestroy(graph); cudaGetLastError(); }
            fg.p = -1;
            throw;
        }
        if (use_graph) {
            CUDA_CHECK(cudaStreamEndCapture(s, &graph));
            const cudaError_t irc = cudaGraphInstantiate(&fg.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (irc != cudaSuccess) { fg.exec = nullptr; fg.p = -1; CUDA_CHECK(irc); }
            fg.kernels = g_launch_count - count0;
            g_launch_count = count0;          // counted when the graph actually runs
        }
    }
    if (use_graph) { CUDA_CHECK(cudaGraphLaunch(fg.exec, s)); g_launch_count += fg.kernels; }
    int h = 0;
    CUDA_CHECK(cudaMemcpyAsync(&h, fg.info.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (info_host) *info_host = h;
    if (h != 0) throw CodeError(B200ADMM_ENOTSPD, "Cholesky factorisation met a non-positive pivot at column " + std::to_string(h));
}
template void spd_inverse<float>(cudaStream_t, float*, i64, i64, float*, int*, float*);
template void spd_inverse<double>(cudaStream_t, double*, i64, i64, double*, int*, double*);

void spd_inverse_f32(cudaStream_t s, float* a, i64 p, i64 ld, float* W, int* info_host)
{
    spd_inverse<float>(s, a, p, ld, W, info_host, nullptr);
}

}  // namespace b200

using namespace b200;

// ---------------------------------------------------------------------------------------------
// exception fence
// ---------------------------------------------------------------------------------------------
template <class F> static int fenced(F&& f)
{
    try {
        f();
        return B200ADMM_OK;
    } catch (const CodeError& e) {
        g_last_error = e.what();
        return e.code;
    } catch (const ArgError& e) {
        g_last_error = e.what();
        return B200ADMM_EINVAL;
    } catch (const CudaError& e) {
        g_last_error = e.what();
        cudaGetLastError();
        return B200ADMM_ECUDA;
    } catch (const std::bad_alloc&) {
        g_last_error = "out of host memory";
        return B200ADMM_ENOMEM;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return B200ADMM_EINVAL;
    }
}

static void check_common(const b200admm_data* d, const b200admm_opts* o)
{
    if (!d || !o) throw ArgError("null argument");
    if (!d->x || !d->y) throw ArgError("x and y must not be null");
    if (d->n <= 0 || d->p <= 0) throw ArgError("x must have positive dimensions");
    if (d->n >= 2147483647LL || d->p >= 2147483647LL) throw ArgError("dimension too large");
    if (o->maxit <= 0) throw ArgError("maxit should be positive");          // R/30_admm_lasso.R:119-124
    if (o->eps_abs < 0 || o->eps_rel < 0) throw ArgError("eps_abs and eps_rel should be nonnegative");
}

extern "C" {

const char* b200admm_last_error(void) { return g_last_error.c_str(); }
int b200admm_version(void) { return B200ADMM_VERSION; }
unsigned long long b200admm_launch_count(void) { return g_launch_count; }
double b200admm_last_gram_seconds(void) { return b200::g_last_gram_seconds; }
void b200admm_release_cache(void) { dev_cache_release(); }
void b200admm_last_work(double* out4) { if (out4) for (int i = 0; i < 4; i++) out4[i] = b200::g_last_work[i]; }
void* b200admm_stream(void)
{
    void* h = nullptr;
    fenced([&] { h = (void*)ctx().stream; });
    return h;
}

int b200admm_device_info(char* name, int name_len, int* sms, int64_t* mem_bytes)
{
    return fenced([&] {
        Context& c = ctx();
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, c.device));
        if (name && name_len > 0) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
        if (sms) *sms = prop.multiProcessorCount;
        if (mem_bytes) *mem_bytes = (int64_t)prop.totalGlobalMem;
    });
}

void b200admm_set_trace(double* buf, int cap, int which, int* nrows)
{
    TraceRequest& t = trace_request();
    t.buf = buf; t.cap = cap; t.which = which; t.nrows = nrows;
    if (nrows) *nrows = 0;
}

void b200admm_set_capture(float* gram_host, float* xy_host, float* stats_host)
{
    CaptureRequest& c = capture_request();
    c.gram = gram_host; c.xy = xy_host; c.stats = stats_host;
}

int b200admm_lasso(const b200admm_data* d, const double* lambda_given, int nlambda_given, int nlambda,
                   double lmin_ratio, int standardize, int intercept, const b200admm_opts* opts, b200admm_path* out)
{
    return fenced([&] {
        check_common(d, opts);
        if (!out) throw ArgError("null result");
        memset(out, 0, sizeof(*out));
        LassoRequest rq;
        rq.d = d; rq.lambda_given = lambda_given; rq.nlambda_given = nlambda_given; rq.nlambda = nlambda;
        rq.lmin_ratio = lmin_ratio; rq.standardize = standardize != 0; rq.intercept = intercept != 0;
        rq.enet = false; rq.alpha = 1.0; rq.opts = *opts;
        try { solve_lasso_like(rq, out); } catch (...) { free_path(out); throw; }
    });
}

int b200admm_enet(const b200admm_data* d, const double* lambda_given, int nlambda_given, int nlambda,
                  double lmin_ratio, int standardize, int intercept, double alpha,
                  const b200admm_opts* opts, b200admm_path* out)
{
    return fenced([&] {
        check_common(d, opts);
        if (!out) throw ArgError("null result");
        if (!(alpha >= 0.0 && alpha <= 1.0)) throw ArgError("alpha must be within [0, 1]");   // R/40_admm_enet.R:31-32
        memset(out, 0, sizeof(*out));
        LassoRequest rq;
        rq.d = d; rq.lambda_given = lambda_given; rq.nlambda_given = nlambda_given; rq.nlambda = nlambda;
        rq.lmin_ratio = lmin_ratio; rq.standardize = standardize != 0; rq.intercept = intercept != 0;
        rq.enet = true; rq.alpha = alpha; rq.opts = *opts;
        try { solve_lasso_like(rq, out); } catch (...) { free_path(out); throw; }
    });
}

int b200admm_parlasso(const b200admm_data* d, const double* lambda_given, int nlambda_given, int nlambda,
                      double lmin_ratio, int standardize, int intercept, int nthread,
                      const b200admm_opts* opts, b200admm_path* out)
{
    return fenced([&] {
        check_common(d, opts);
        if (!out) throw ArgError("null result");
        if (nthread < 1) throw ArgError("nthread must be a positive integer");              // R/30_admm_lasso.R:105-106
        memset(out, 0, sizeof(*out));
        LassoRequest rq;
        rq.d = d; rq.lambda_given = lambda_given; rq.nlambda_given = nlambda_given; rq.nlambda = nlambda;
        rq.lmin_ratio = lmin_ratio; rq.standardize = standardize != 0; rq.intercept = intercept != 0;
        rq.enet = false; rq.alpha = 1.0; rq.opts = *opts;
        try { solve_consensus(rq, nthread, out); } catch (...) { free_path(out); throw; }
    });
}

void b200admm_free_path(b200admm_path* out) { free_path(out); }

int b200admm_lad(const b200admm_data* d, int intercept, const b200admm_opts* opts, b200admm_dense* out)
{
    return fenced([&] {
        check_common(d, opts);
        if (!out) throw ArgError("null result");
        memset(out, 0, sizeof(*out));
        if (d->n <= d->p) throw ArgError("nrow(x) must be greater than ncol(x)");            // R/20_admm_lad.R:21-22
        try { solve_lad(d, intercept != 0, *opts, out); } catch (...) { free(out->beta); out->beta = nullptr; throw; }
    });
}
void b200admm_free_dense(b200admm_dense* out)
{
    if (!out) return;
    free(out->beta);
    out->beta = nullptr;
}

int b200admm_bp(const b200admm_data* d, const b200admm_opts* opts, b200admm_path* out)
{
    return fenced([&] {
        check_common(d, opts);
        if (!out) throw ArgError("null result");
        memset(out, 0, sizeof(*out));
        if (d->p <= d->n) throw ArgError("ncol(x) must be greater than nrow(x)");            // R/10_admm_bp.R:30-31
        try { solve_bp(d, *opts, out); } catch (...) { free_path(out); throw; }
    });
}

int b200admm_synth_f32(void* x_dev, void* y_dev, int64_t nrows, int64_t p, int64_t row0, uint64_t seed,
                       float mean_x, float sd_x, int nsig, float noise)
{
    return fenced([&] {
        Context& c = ctx();
        if (!x_dev || nrows <= 0 || p <= 0) throw ArgError("synth: bad arguments");
        synth_design_f32(c.stream, (float*)x_dev, (float*)y_dev, nrows, p, row0, seed, mean_x, sd_x, nsig, noise);
        CUDA_CHECK(cudaStreamSynchronize(c.stream));
    });
}

int b200admm_comm_id(void* id_out)
{
    return fenced([&] {
        if (!id_out) throw ArgError("null id");
        ctx();
        comm_unique_id(id_out);
    });
}
int b200admm_comm_init(const void* id, int rank, int nranks)
{
    return fenced([&] {
        if (!id) throw ArgError("null id");
        ctx();
        comm_init(id, rank, nranks);
    });
}
void b200admm_comm_destroy(void)
{
    try { comm_destroy(); } catch (...) {}
}
void b200admm_comm_suspend(int on) { comm().suspended = on != 0; }

// ---- kernel-level entry points ----------------------------------------------------------------
int b200admm_k_standardize_f32(const void* x_in, void* x_out, void* y_inout, int64_t n, int64_t p,
                               int standardize, int intercept, float* meanx_host, float* scalex_host,
                               float* meany_scaley_host)
{
    return fenced([&] {
        Context& c = ctx();
        const int flag = (standardize ? 1 : 0) + (intercept ? 2 : 0);
        DevBuf<float> meanX(p), scaleX(p), tmp(2 * p + 4), y2(2);
        meanX.zero(c.stream);
        CUDA_CHECK(cudaMemsetAsync(y2.p, 0, 2 * sizeof(float), c.stream));
        if (x_in) standardize_columns<float>(c.stream, (const float*)x_in, (float*)x_out, n, p, n, flag, meanX.p, scaleX.p, tmp.p);
        if (y_inout) standardize_y<float>(c.stream, (float*)y_inout, n, flag, y2.p, tmp.p);
        if (meanx_host) CUDA_CHECK(cudaMemcpyAsync(meanx_host, meanX.p, p * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
        if (scalex_host) CUDA_CHECK(cudaMemcpyAsync(scalex_host, scaleX.p, p * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
        if (meany_scaley_host) CUDA_CHECK(cudaMemcpyAsync(meany_scaley_host, y2.p, 2 * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
        CUDA_CHECK(cudaStreamSynchronize(c.stream));
    });
}

int b200admm_k_gram_f32(const void* x, int64_t n, int64_t p, void* g, int use_tensor)
{
    return fenced([&] {
        Context& c = ctx();
        bool done = false;
        if (use_tensor) {
            CUDA_CHECK(cudaMemsetAsync(g, 0, sizeof(float) * (size_t)p * (size_t)p, c.stream));
            // use_tensor: 1 = TF32 truncation split, 2 = TF32 round-to-nearest split, 3 = fp16 split (unit-scale data)
            const int split = use_tensor >= 3 ? GRAM_SPLIT_F16 : (use_tensor == 2 ? GRAM_SPLIT_TF32 : GRAM_SPLIT_TRUNC);
            done = gram_tn_tensor(c.stream, (const float*)x, n, n, p, (float*)g, p, split);
            if (!done) throw ArgError("tensor-core Gram kernel cannot take this shape (needs n % 4 == 0 and p >= 8)");
            if (split == GRAM_SPLIT_F16 && gram_f16_overflowed(c.stream))
                throw ArgError("fp16-split Gram kernel: a value exceeds the fp16 range (use mode 2)");
        }
        if (!done)
            gemm<float>(c.stream, true, false, p, p, n, 1.f, (const float*)x, n, (const float*)x, n, 0.f, (float*)g, p, GEMM_LOWER | GEMM_MIRROR);
        CUDA_CHECK(cudaStreamSynchronize(c.stream));
    });
}

int b200admm_k_gram_plan(int ntiles, int npairs, int nk, int* cover, long long* per_pair, int* nslices, int* split_tiles)
{
    return fenced([&] {
        if (ntiles < 1 || npairs < 1 || nk < 1 || !cover || !per_pair) throw ArgError("gram plan: bad arguments");
        gram_f16_plan(ntiles, npairs, nk, cover, per_pair, nslices, split_tiles);
    });
}

int b200admm_k_panel_schedule(int64_t p, int64_t panel_cols, int64_t* begin, int cap)
{
    try {
        if (p < 1 || !begin || cap < 2) return -1;
        const std::vector<i64> b = host_panel_schedule(p, panel_cols);
        if ((int)b.size() > cap) return -1;
        for (size_t i = 0; i < b.size(); i++) begin[i] = b[i];
        return (int)b.size() - 1;
    } catch (...) { return -1; }
}

int b200admm_k_tri_plan(int p, int sms, int* rows, int cap, long long* smem_bytes)
{
    try {
        if (p < 1 || sms < 1) return -1;
        return tall_tri_plan(p, sms, rows, cap, smem_bytes);
    } catch (...) { return -1; }
}

int b200admm_k_lambda_grid(double lmax, double lmin_ratio, int nlambda, double* out)
{
    return fenced([&] {
        if (!(lmax > 0) || !(lmin_ratio > 0) || nlambda < 1 || !out) throw ArgError("lambda grid: bad arguments");
        std::vector<double> g;
        make_lambda_grid(lmax, lmin_ratio, nlambda, g);
        for (int i = 0; i < nlambda; i++) out[i] = g[i];
    });
}

int b200admm_k_gemv_t_f32(const void* a, int64_t m, int64_t ncol, const void* v, void* out)
{
    return fenced([&] {
        Context& c = ctx();
        gemv_t<float>(c.stream, (const float*)a, m, ncol, m, (const float*)v, (float*)out);
        CUDA_CHECK(cudaStreamSynchronize(c.stream));
    });
}

int b200admm_k_coarse_eig_f32(const void* sm, int64_t n, float* ev_host, int* info_host)
{
    return fenced([&] {
        Context& c = ctx();
        if (!sm || !ev_host) throw ArgError("coarse eig: null argument");
        if (n < 3) throw CodeError(B200ADMM_ELANCZOS, "coarse eigenvalue estimate needs at least 3 variables (Spectra: 1 <= nev < ncv <= n)");
        DevBuf<float> dv(n), dw(n);
        eig::LanczosInfo info;
        cudaStream_t s = c.stream;
        const float* S = (const float*)sm;
        auto op = [&](const float* v, float* w) {
            CUDA_CHECK(cudaMemcpyAsync(dv.p, v, n * sizeof(float), cudaMemcpyHostToDevice, s));
            gemv_t<float>(s, S, n, n, n, dv.p, dw.p);
            CUDA_CHECK(cudaMemcpyAsync(w, dw.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaStreamSynchronize(s));
        };
        *ev_host = eig::coarse_largest_eigenvalue<float>(op, n, &info);
        if (info_host) { info_host[0] = info.nmatvec; info_host[1] = info.nrestart; info_host[2] = info.converged; }
    });
}

int b200admm_k_gemm_tn_f32(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                           void* c, int64_t ldc, int tile_mode, int klo_mode, int khi_mode, int epi)
{
    return fenced([&] {
        Context& cx = ctx();
        if (!gemm_tn_tensor(cx.stream, (const float*)a, lda, (const float*)b, ldb, m, n, k, (float*)c, ldc, tile_mode, klo_mode, khi_mode, epi))
            throw ArgError("tensor-core TN product cannot take this shape (leading dimensions % 4, 16-byte aligned bases, even SM count)");
        CUDA_CHECK(cudaStreamSynchronize(cx.stream));
    });
}

int b200admm_k_gemm_f64(int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha, const void* a, int64_t lda,
                        const void* b, int64_t ldb, double beta, void* c, int64_t ldc, int mode, float* ms_out, int repeats)
{
    return fenced([&] {
        Context& cx = ctx();
        if (m < 1 || n < 1 || k < 1 || !a || !b || !c) throw ArgError("gemm_f64: bad arguments");
        const int reps = repeats < 1 ? 1 : repeats;
        EventTimer t(cx.stream);
        t.start();
        for (int r = 0; r < reps; r++)
            gemm<double>(cx.stream, ta != 0, tb != 0, m, n, k, alpha, (const double*)a, lda, (const double*)b, ldb, beta, (double*)c, ldc, mode);
        const double sec = t.stop();
        if (ms_out) *ms_out = (float)(sec * 1e3 / reps);
    });
}

int b200admm_k_chol_f32(void* a, int64_t p, int* info_host)
{
    return fenced([&] {
        Context& c = ctx();
        DevBuf<float> work(chol_work<float>(p));
        DevBuf<int> info(1);
        chol_lower<float>(c.stream, (float*)a, p, p, work.p, info.p);
        int h = 0;
        CUDA_CHECK(cudaMemcpyAsync(&h, info.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        CUDA_CHECK(cudaStreamSynchronize(c.stream));
        if (info_host) *info_host = h;
    });
}

// debug / tuning aid, not part of the documented ABI: mean ms per launch of a one-CTA diagonal-block kernel
extern "C" double b200admm_debug_diag_ms(int mode, int reps, void* a, int64_t lda, void* dinv)
{
    double ms = -1;
    fenced([&] {
        Context& c = ctx();
        DevBuf<int> info(1);
        info.zero(c.stream);
        ms = diag_kernel_bench(c.stream, mode, reps, (float*)a, lda, (float*)dinv, info.p);
    });
    return ms;
}

int b200admm_k_spd_inverse_f32(void* a, int64_t p, void* work, int* info_host)
{
    return fenced([&] {
        Context& c = ctx();
        spd_inverse_f32(c.stream, (float*)a, p, p, (float*)work, info_host);
        CUDA_CHECK(cudaStreamSynchronize(c.stream));
    });
}

int b200admm_k_fused_zu_f32(const void* x, const void* adj_y, const void* old_z, const void* adj_z,
                            void* z, void* y, int64_t len, double lambda, double rho, int enet, double alpha,
                            double* sums_host, float* ms_out, int repeats)
{
    return fenced([&] {
        Context& c = ctx();
        DevBuf<double> sums(6);
        DevBuf<float> part((size_t)6 * fused_zu_blocks());
        if (repeats < 1) repeats = 1;
        EventTimer t(c.stream);
        // one untimed pass, then `repeats` timed passes
        fused_zu_pass(c.stream, (const float*)x, (const float*)adj_y, (const float*)old_z, (const float*)adj_z,
                      (float*)z, (float*)y, len, lambda, rho, enet, alpha, sums.p, part.p);
        t.start();
        for (int r = 0; r < repeats; r++)
            fused_zu_pass(c.stream, (const float*)x, (const float*)adj_y, (const float*)old_z, (const float*)adj_z,
                          (float*)z, (float*)y, len, lambda, rho, enet, alpha, sums.p, part.p);
        const double sec = t.stop();
        if (ms_out) *ms_out = (float)(sec * 1e3 / repeats);
        if (sums_host) CUDA_CHECK(cudaMemcpy(sums_host, sums.p, 6 * sizeof(double), cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"
