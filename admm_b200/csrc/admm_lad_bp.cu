// admm_lad_bp.cu -- least absolute deviation and basis pursuit (float64, accelerated ADMM).
//
// Reference being replaced (all in /root/reference/src):
//   admm_lad()  LAD.cpp:16-48      ADMMLAD  ADMMLAD.h:62-225   (x-update = projection onto Range(X))
//   admm_bp()   BP.cpp:20-46       ADMMBP   ADMMBP.h:48-197    (x-update = projection onto {Ax = b})
//   FADMMBase::solve / update_rho   FADMMBase.h:109-133,185-265
//
// Both solvers spend an iteration in two matrix-vector passes over an n x p float64 matrix
// (X for LAD, M = L^-1 A for BP: 2 * 8 n p bytes, HBM-bound) plus one fused vector kernel that
// does the prox, the residual, the dual update and all six squared norms in a single pass
// (z + u + residual kernel, K8 of SURVEY.md).  An iteration is milliseconds at the benchmark
// sizes, so the loop is driven from the host: one 48-byte device->host read of the norms per
// iteration decides convergence, the Nesterov restart and the rho balancing exactly as the
// reference does, in double.
//
// Compiled with --fmad=false: the element-wise expressions are evaluated unfused, in the
// reference's order ((y - adj_y / rho) + adj_z,  (x - y) + adj_y / rho, ...).
#include "solvers.h"
#include "kernels.h"
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace b200 {

namespace {

constexpr int VT = 256;
constexpr int NSUM = 6;

// v = (c - adj_y / rho) + adj_z      (c == nullptr: v = -adj_y / rho + adj_z)
__global__ void __launch_bounds__(VT) prep_v_kernel(const double* __restrict__ c, const double* __restrict__ adj_y,
                                                    const double* __restrict__ adj_z, double rho, i64 len, double* __restrict__ v)
{
    const i64 stride = (i64)gridDim.x * VT;
    for (i64 i = (i64)blockIdx.x * VT + threadIdx.x; i < len; i += stride) {
        double t = c ? (c[i] - adj_y[i] / rho) : (-adj_y[i] / rho);
        const double az = adj_z[i];
        if (az != 0.0) t += az;
        v[i] = t;
    }
}

// x = v + q - r     (BP, ADMMBP.h:57,66)
__global__ void __launch_bounds__(VT) bp_combine_kernel(const double* __restrict__ v, const double* __restrict__ q,
                                                        const double* __restrict__ r, i64 len, double* __restrict__ x)
{
    const i64 stride = (i64)gridDim.x * VT;
    for (i64 i = (i64)blockIdx.x * VT + threadIdx.x; i < len; i += stride) {
        const double a = v[i] + q[i];
        x[i] = -1.0 * r[i] + a;
    }
}

// fused z + u + residual + norms.  shift == y_data for LAD (works on x - y), nullptr for BP.
__global__ void __launch_bounds__(VT) fused_step_kernel(const double* __restrict__ x, const double* __restrict__ shift,
                                                        const double* __restrict__ adj_y, const double* __restrict__ old_z,
                                                        const double* __restrict__ adj_z, double rho, i64 len,
                                                        double* __restrict__ z, double* __restrict__ y, double* __restrict__ part)
{
    const double pen = 1.0 / rho;
    double ps[NSUM] = {0, 0, 0, 0, 0, 0};
    const i64 stride = (i64)gridDim.x * VT;
    for (i64 i = (i64)blockIdx.x * VT + threadIdx.x; i < len; i += stride) {
        const double xv = x[i];
        const double xs = shift ? xv - shift[i] : xv;
        const double ay = adj_y[i];
        const double u = xs + ay / rho;
        double zn;
        if (u > pen) zn = u - pen; else if (u < -pen) zn = u + pen; else zn = 0.0;
        const double res = xs - zn;
        const double yn = ay + rho * res;
        z[i] = zn;
        y[i] = yn;
        const double d1 = zn - old_z[i], d2 = zn - adj_z[i];
        ps[0] += res * res; ps[1] += d1 * d1; ps[2] += d2 * d2;
        ps[3] += xv * xv;   ps[4] += zn * zn; ps[5] += yn * yn;
    }
    __shared__ double s_red[VT / 32][NSUM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NSUM; q++) ps[q] = warp_sum(ps[q]);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NSUM; q++) s_red[warp][q] = ps[q];
    }
    __syncthreads();
    if (threadIdx.x < NSUM) {
        double s = 0.0;
        for (int w = 0; w < VT / 32; w++) s += s_red[w][threadIdx.x];
        part[(size_t)blockIdx.x * NSUM + threadIdx.x] = s;
    }
}
__global__ void finish_sums_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ sums)
{
    const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (q >= NSUM) return;
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += part[(size_t)b * NSUM + q];
    s = warp_sum(s);
    if (lane == 0) sums[q] = s;
}

// acceleration (adj = (1+t) new - t old) or restart (adj = old)   FADMMBase.h:243-256
__global__ void __launch_bounds__(VT) accel_kernel(const double* __restrict__ z_new, const double* __restrict__ z_old,
                                                   const double* __restrict__ y_new, const double* __restrict__ y_old,
                                                   int accel, double c1, double c2, i64 len,
                                                   double* __restrict__ adj_z, double* __restrict__ adj_y)
{
    const i64 stride = (i64)gridDim.x * VT;
    for (i64 i = (i64)blockIdx.x * VT + threadIdx.x; i < len; i += stride) {
        if (accel) {
            adj_z[i] = c1 * z_new[i] - c2 * z_old[i];
            adj_y[i] = c1 * y_new[i] - c2 * y_old[i];
        } else {
            adj_z[i] = z_old[i];
            adj_y[i] = y_old[i];
        }
    }
}

__global__ void sumsq_kernel(const double* __restrict__ a, i64 len, double* out)
{
    __shared__ double scratch[33];
    double s = 0.0;
    for (i64 i = threadIdx.x; i < len; i += blockDim.x) s += a[i] * a[i];
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) *out = s;
}

inline unsigned vgrid(i64 len) { return (unsigned)std::max<i64>(1, std::min<i64>((len + VT - 1) / VT, (i64)sm_count() * 8)); }

// rho balancing (FADMMBase.h:109-133)
inline void balance_rho(double& rho, double rp, double ep, double rd, double ed)
{
    if (rp / ep > 10 * rd / ed) rho *= 2;
    else if (rd / ed > 10 * rp / ep) rho /= 2;
    if (rp < ep) rho /= 1.2;
    if (rd < ed) rho *= 1.2;
}

// The accelerated loop shared by LAD and BP.  XStep(v, x) computes the model's x-update from v.
struct FadmmF64 {
    cudaStream_t s;
    i64 dim;
    const double* shift;        // y_data (LAD) or nullptr (BP)
    double extra_norm;          // ||y_data|| for LAD's eps_primal, 0 for BP
    double eps_abs, eps_rel;
    double rho;
    DevBuf<double> x, v, zbuf[2], ybuf[2], adj_z, adj_y, part, sums;
    int nblk;

    FadmmF64(cudaStream_t s_, i64 dim_, const double* shift_, double extra, double ea, double er, double rho_)
        : s(s_), dim(dim_), shift(shift_), extra_norm(extra), eps_abs(ea), eps_rel(er), rho(rho_)
    {
        x.alloc(dim); v.alloc(dim); adj_z.alloc(dim); adj_y.alloc(dim);
        for (int i = 0; i < 2; i++) { zbuf[i].alloc(dim); ybuf[i].alloc(dim); zbuf[i].zero(s); ybuf[i].zero(s); }
        x.zero(s); adj_z.zero(s); adj_y.zero(s);
        nblk = (int)vgrid(dim);
        part.alloc((size_t)nblk * NSUM);
        sums.alloc(NSUM);
    }
    void prep_v()
    {
        prep_v_kernel<<<vgrid(dim), VT, 0, s>>>(shift, adj_y.p, adj_z.p, rho, dim, v.p);
        KERNEL_CHECK();
    }
    template <class XStep> int solve(int maxit, XStep&& xstep, int& cur_out)
    {
        TraceRequest& tr = trace_request();
        const bool tracing = tr.buf && tr.cap > 0 && tr.which == 0;
        double sx2 = 0, sz2 = 0, sy2 = 0;
        double adj_a = 1.0, adj_c = 9999.0;
        const double sqd = std::sqrt((double)dim);
        int cur = 0, i;
        prep_v();
        for (i = 0; i < maxit; i++) {
            const int nxt = cur ^ 1;
            double r = std::max(std::sqrt(sx2), std::sqrt(sz2));
            r = std::max(r, extra_norm);
            const double eps_primal = r * eps_rel + sqd * eps_abs;
            const double eps_dual = std::sqrt(sy2) * eps_rel + sqd * eps_abs;
            xstep(v.p, x.p);
            fused_step_kernel<<<nblk, VT, 0, s>>>(x.p, shift, adj_y.p, zbuf[cur].p, adj_z.p, rho, dim, zbuf[nxt].p, ybuf[nxt].p, part.p);
            KERNEL_CHECK();
            finish_sums_kernel<<<1, 32 * NSUM, 0, s>>>(part.p, nblk, sums.p);
            KERNEL_CHECK();
            double h[NSUM];
            CUDA_CHECK(cudaMemcpyAsync(h, sums.p, sizeof h, cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaStreamSynchronize(s));
            const double resid_primal = std::sqrt(h[0]);
            const double resid_dual = rho * std::sqrt(h[1]);
            sx2 = h[3]; sz2 = h[4]; sy2 = h[5];
            if (tracing && i < tr.cap) {
                double* row = tr.buf + 5 * (size_t)i;
                row[0] = eps_primal; row[1] = resid_primal; row[2] = eps_dual; row[3] = resid_dual; row[4] = rho;
                if (tr.nrows) *tr.nrows = i + 1;
            }
            const int old = cur;
            cur = nxt;
            if (resid_primal < eps_primal && resid_dual < eps_dual) break;
            const double old_c = adj_c;
            adj_c = rho * resid_primal * resid_primal + rho * h[2];
            int accel = 0;
            double c1 = 0, c2 = 0;
            if (adj_c < 0.999 * old_c) {
                const double old_a = adj_a;
                adj_a = 0.5 + 0.5 * std::sqrt(1 + 4.0 * old_a * old_a);
                const double ratio = (old_a - 1.0) / adj_a;
                c1 = 1 + ratio; c2 = ratio; accel = 1;
            } else {
                adj_a = 1.0;
                adj_c = old_c / 0.999;
            }
            accel_kernel<<<vgrid(dim), VT, 0, s>>>(zbuf[cur].p, zbuf[old].p, ybuf[cur].p, ybuf[old].p, accel, c1, c2, dim, adj_z.p, adj_y.p);
            KERNEL_CHECK();
            if (i > 5) balance_rho(rho, resid_primal, eps_primal, resid_dual, eps_dual);
            prep_v();                                     // v for the next x-update, with the balanced rho
        }
        cur_out = cur;
        return i + 1;
    }
};

double device_norm(cudaStream_t s, const double* a, i64 len)
{
    DevBuf<double> d(1);
    sumsq_kernel<<<1, 1024, 0, s>>>(a, len, d.p);
    KERNEL_CHECK();
    double h = 0;
    CUDA_CHECK(cudaMemcpyAsync(&h, d.p, sizeof h, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    return std::sqrt(h);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// admm_lad  (LAD.cpp:16-48)
// ---------------------------------------------------------------------------------------------
void solve_lad(const b200admm_data* d, bool intercept, const b200admm_opts& o, b200admm_dense* out)
{
    Context& c = ctx();
    cudaStream_t s = c.stream;
    const i64 n = d->n, p = d->p;
    if (d->dtype != B200ADMM_F64_HOST && d->dtype != B200ADMM_F64_DEVICE)
        throw ArgError("admm_lad computes in float64: pass float64 data");
    const double t_begin = wall_now();
    EventTimer tm(s);
    b200admm_timing T;
    memset(&T, 0, sizeof T);
    const int flag = 1 + (intercept ? 2 : 0);                       // DataStd<double>(n, p, true, intercept)

    DevBuf<double> X((size_t)n * (size_t)p), Y(n);
    tm.start();
    ingest_f64(s, d->x, d->dtype, (size_t)n * (size_t)p, X.p);
    ingest_f64(s, d->y, d->dtype, (size_t)n, Y.p);
    T.ingest = tm.stop();

    DevBuf<double> d_meanX(p), d_scaleX(p), tmp(2 * p + 8), y2(2);
    tm.start();
    d_meanX.zero(s);
    CUDA_CHECK(cudaMemsetAsync(y2.p, 0, 2 * sizeof(double), s));
    standardize_y<double>(s, Y.p, n, flag, y2.p, tmp.p);
    standardize_columns<double>(s, X.p, X.p, n, p, n, flag, d_meanX.p, d_scaleX.p, tmp.p);
    std::vector<double> meanX(p, 0.0), scaleX(p, 1.0);
    double h2[2] = {0, 1};
    if (flag == 3) CUDA_CHECK(cudaMemcpyAsync(meanX.data(), d_meanX.p, p * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(scaleX.data(), d_scaleX.p, p * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(h2, y2.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
    T.standardize = tm.stop();
    const double meanY = flag == 3 ? h2[0] : 0.0, scaleY = h2[1];

    // X'X, factor, explicit inverse (replaces LLT::compute + solve, ADMMLAD.h:186-189,76)
    DevBuf<double> Ginv((size_t)p * (size_t)p), W((size_t)p * (size_t)p);
    tm.start();
    gemm<double>(s, true, false, p, p, n, 1.0, X.p, n, X.p, n, 0.0, Ginv.p, p, GEMM_LOWER | GEMM_MIRROR);
    T.gram = tm.stop();
    tm.start();
    int info = 0;
    spd_inverse<double>(s, Ginv.p, p, p, W.p, &info, nullptr);
    DevBuf<double> H;
    const bool use_hat = n <= 2000;                                  // ADMMLAD.h:66,191
    if (use_hat) {
        DevBuf<double> Tm((size_t)n * (size_t)p);
        // T = X L^-T = X W'   ;   H = T T'
        gemm<double>(s, false, true, n, p, p, 1.0, X.p, n, W.p, p, 0.0, Tm.p, n, GEMM_BT_LOWER_TRI);
        H.alloc((size_t)n * (size_t)n);
        gemm<double>(s, false, true, n, n, p, 1.0, Tm.p, n, Tm.p, n, 0.0, H.p, n, GEMM_LOWER | GEMM_MIRROR);
        CUDA_CHECK(cudaStreamSynchronize(s));
    }
    W.release();
    T.factor = tm.stop();

    const double ynorm = device_norm(s, Y.p, n);
    FadmmF64 F(s, n, Y.p, ynorm, o.eps_abs, o.eps_rel, o.rho);
    DevBuf<double> t1(p), t2(p), work(gemv_n_work(n, p));
    auto xstep = [&](const double* v, double* x) {
        if (use_hat) {
            gemv_t<double>(s, H.p, n, n, n, v, x);                   // H symmetric: H'v == Hv  (dsymv_)
        } else {
            gemv_t<double>(s, X.p, n, p, n, v, t1.p);                // X'v
            gemv_t<double>(s, Ginv.p, p, p, p, t1.p, t2.p);          // (X'X)^-1 (.)
            gemv_n<double>(s, X.p, n, p, n, t2.p, x, work.p);        // X (.)
        }
    };
    tm.start();
    int cur = 0;
    const int niter = F.solve(o.maxit, xstep, cur);
    T.iterate = tm.stop();

    // get_x(): beta = (X'X)^-1 X' (y - adj_y / rho + adj_z)  with the final adj_* and rho (ADMMLAD.h:220-225)
    tm.start();
    F.prep_v();
    gemv_t<double>(s, X.p, n, p, n, F.v.p, t1.p);
    gemv_t<double>(s, Ginv.p, p, p, p, t1.p, t2.p);
    std::vector<double> coef(p);
    CUDA_CHECK(cudaMemcpyAsync(coef.data(), t2.p, p * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    // dense recover (DataStd.h:159-181)
    out->beta = (double*)malloc(sizeof(double) * (p + 1));
    if (!out->beta) throw CodeError(B200ADMM_ENOMEM, "out of host memory");
    double acc = 0;
    for (i64 j = 0; j < p; j++) {
        coef[j] /= scaleX[j];
        coef[j] *= scaleY;
        if (flag == 3) acc += coef[j] * meanX[j];
        out->beta[j + 1] = coef[j];
    }
    out->beta[0] = flag == 3 ? meanY - acc : 0.0;
    out->len = p + 1;
    out->niter = niter;
    out->rho = F.rho;
    T.finish = tm.stop();
    T.total = wall_now() - t_begin;
    out->t = T;
}

// ---------------------------------------------------------------------------------------------
// admm_bp  (BP.cpp:20-46)
// ---------------------------------------------------------------------------------------------
void solve_bp(const b200admm_data* d, const b200admm_opts& o, b200admm_path* out)
{
    Context& c = ctx();
    cudaStream_t s = c.stream;
    const i64 n = d->n, p = d->p;
    if (d->dtype != B200ADMM_F64_HOST && d->dtype != B200ADMM_F64_DEVICE)
        throw ArgError("admm_bp computes in float64: pass float64 data");
    const double t_begin = wall_now();
    EventTimer tm(s);
    b200admm_timing T;
    memset(&T, 0, sizeof T);

    DevBuf<double> A((size_t)n * (size_t)p), b(n);
    tm.start();
    ingest_f64(s, d->x, d->dtype, (size_t)n * (size_t)p, A.p);
    ingest_f64(s, d->y, d->dtype, (size_t)n, b.p);
    T.ingest = tm.stop();

    // AA' (n x n), its inverse and W = L^-1;  q = A'(AA')^-1 b;  M = L^-1 A   (ADMMBP.h:165-182)
    DevBuf<double> K((size_t)n * (size_t)n), W((size_t)n * (size_t)n), M((size_t)n * (size_t)p), q(p), t1(n);
    tm.start();
    gemm<double>(s, false, true, n, n, p, 1.0, A.p, n, A.p, n, 0.0, K.p, n, GEMM_LOWER | GEMM_MIRROR);
    T.gram = tm.stop();
    tm.start();
    int info = 0;
    spd_inverse<double>(s, K.p, n, n, W.p, &info, nullptr);
    gemv_t<double>(s, K.p, n, n, n, b.p, t1.p);                      // (AA')^-1 b
    gemv_t<double>(s, A.p, n, p, n, t1.p, q.p);                      // A' (.)
    gemm<double>(s, false, false, n, p, n, 1.0, W.p, n, A.p, n, 0.0, M.p, n, GEMM_A_LOWER_TRI);
    CUDA_CHECK(cudaStreamSynchronize(s));
    A.release(); K.release(); W.release();
    T.factor = tm.stop();

    FadmmF64 F(s, p, nullptr, 0.0, o.eps_abs, o.eps_rel, o.rho);
    DevBuf<double> wk(n), r(p), work(gemv_n_work(n, p));
    auto xstep = [&](const double* v, double* x) {
        gemv_n<double>(s, M.p, n, p, n, v, wk.p, work.p);            // M v
        gemv_t<double>(s, M.p, n, p, n, wk.p, r.p);                  // M'(M v)
        bp_combine_kernel<<<vgrid(p), VT, 0, s>>>(v, q.p, r.p, p, x);
        KERNEL_CHECK();
    };
    tm.start();
    int cur = 0;
    const int niter = F.solve(o.maxit, xstep, cur);
    T.iterate = tm.stop();

    tm.start();
    std::vector<std::vector<double>> cols(1, std::vector<double>(p));
    CUDA_CHECK(cudaMemcpyAsync(cols[0].data(), F.zbuf[cur].p, p * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    out->nlambda = 1;
    out->lambda = (double*)malloc(sizeof(double));
    out->niter = (int*)malloc(sizeof(int));
    if (!out->lambda || !out->niter) throw CodeError(B200ADMM_ENOMEM, "out of host memory");
    out->lambda[0] = 0.0;
    out->niter[0] = niter;
    std::vector<double> none;
    assemble_csc<double>(cols, none, false, p, out);
    out->rho = F.rho;
    T.finish = tm.stop();
    T.total = wall_now() - t_begin;
    out->t = T;
}

}  // namespace b200
