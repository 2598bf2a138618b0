// admm_consensus.cu -- admm_parlasso: the reference's row-split (observation-split) global
// consensus lasso, one block per GPU with ONE all-reduce per iteration.
//
// Reference being replaced (all in /root/reference/src):
//   admm_parlasso()                         ParLasso.cpp:33-111
//   PADMMBase_Master::solve / update_x / update_z / update_y / eps   PADMMBase.h:117-237
//   PADMMBase_Worker::update_y                                        PADMMBase.h:65-74
//   PADMMLasso_Worker::next_x / init / add_xu_to (incl. Woodbury)     PADMMLasso.h:17-68
//   PADMMLasso_Master ctor (row split) / next_z / resid_dual / init   PADMMLasso.h:99-212
//
// Block i owns rows [i * floor(n/N), ...) (the last block also takes the remainder) and all p
// columns; it keeps x_i, y_i and K_i^-1 = (A_i'A_i + rho I)^-1 (rho is fixed, so the reference's
// per-iteration LLT solve becomes one bandwidth-bound product with the explicit inverse; blocks
// with fewer rows than columns use the reference's Woodbury form with (A_i A_i' + rho I)^-1).
//
// Exchange step (the only one): the packed vector
//     [ sum_i (x_i + y_i / rho)  (p floats) | sum_i |x_i|^2 | sum_i |x_i - z|^2 (previous iteration) | sum_i |y_i|^2 ]
// is summed over ranks with one ncclAllReduce of p + 3 floats over NVLink.  The result is
// bit-identical on every rank, so the z-update is computed redundantly everywhere and needs no
// broadcast.  The primal residual of iteration t needs z(t), which only exists after the
// exchange; it rides on the payload of iteration t + 1, i.e. the stopping rule is evaluated one
// exchange late (the iterates, the returned z and the iteration count are exactly those of the
// serial master loop; the price is one extra x-update per lambda).
//
// Without a communicator the N blocks live on one device and the same code runs with the
// all-reduce degenerated to a no-op.  Compiled with --fmad=false.
#include "solvers.h"
#include "kernels.h"
#include "comm.h"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <memory>

namespace b200 {

void finish_lasso_path(const std::vector<float>& z_all, int nl, i64 p, int flag, const std::vector<float>& meanX,
                       const std::vector<float>& scaleX, float meanY, float scaleY, b200admm_path* out);
struct StdStats {
    std::vector<float> meanX, scaleX;
    float meanY = 0.f, scaleY = 1.f;
};
void standardize_all(cudaStream_t s, const float* X_in, i64 ld_in, float* X_out, i64 ld_out, float* y, i64 n_local, i64 n_total, i64 p,
                     int flag, float* d_meanX, float* d_scaleX, StdStats& st);
void standardize_y_sharded(cudaStream_t s, float* y, i64 n_local, i64 n_total, int flag, StdStats& st);
void standardize_stats3_sharded(cudaStream_t s, const float* X_in, i64 ld_in, i64 n_local, i64 n_total, i64 pc,
                                float* d_meanX, float* d_scaleX, float* d_inv, float* tmp);
void fetch_std_stats(cudaStream_t s, i64 p, int flag, const float* d_meanX, const float* d_scaleX, StdStats& st);

namespace {

constexpr int CT = 1024;

// rhs = Ab - y + rho * z   (the rho * z term in double, only where z != 0: PADMMLasso.h:19-21)
__global__ void __launch_bounds__(CT) cons_rhs_kernel(const float* __restrict__ Ab, const float* __restrict__ y, const float* __restrict__ z,
                                                      double rho, int p, float* __restrict__ rhs)
{
    for (int j = blockIdx.x * CT + threadIdx.x; j < p; j += gridDim.x * CT) {
        float v = Ab[j] - y[j];
        const float zj = z[j];
        if (zj != 0.f) v = (float)((double)v + rho * (double)zj);
        rhs[j] = v;
    }
}
// Woodbury tail: x = (rhs - t) / frho
__global__ void __launch_bounds__(CT) cons_woodbury_kernel(const float* __restrict__ rhs, const float* t, float frho, int p, float* x)   /* t may alias x (in place) */
{
    for (int j = blockIdx.x * CT + threadIdx.x; j < p; j += gridDim.x * CT) x[j] = (rhs[j] - t[j]) / frho;
}
// acc (+)= x + y / frho ; slot += |x|^2        (single CTA: deterministic)
__global__ void __launch_bounds__(CT) cons_gather_kernel(const float* __restrict__ x, const float* __restrict__ y, float frho, int p,
                                                         int first, float* __restrict__ acc, float* __restrict__ slot_x2)
{
    __shared__ float scratch[33];
    float s = 0.f;
    for (int j = threadIdx.x; j < p; j += CT) {
        const float xv = x[j];
        const float t = xv + y[j] / frho;
        acc[j] = first ? t : acc[j] + t;
        s += xv * xv;
    }
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) *slot_x2 = first ? s : *slot_x2 + s;
}
// Device-side loop control (batched passes): the master's bookkeeping of PADMMBase_Master::solve (src/PADMMBase.h:174-237)
// -- closing the books on iteration t - 1 with the primal residual that arrived with this exchange, the stopping rule,
// the tolerances of iteration t, the dual residual after the z-update -- in the host loop's own double expressions.
struct ConsCtl {
    double sx2, sz2, eps_p, eps_d, rd;     // norms of the current iterate, tolerances / dual residual of the open iteration
    int have_prev, stop, niter, t, trace_rows;
};
struct ConsConst {
    double eps_abs, eps_rel, sqrt_pN, sqrtN, rho, dN;
    int maxit;
};

// z_new = soft(acc / N, pen); out[0] = sum (z_new - z)^2, out[1] = |z_new|^2 ; z <- z_new
// acc = the exchanged payload (p sums + 3 tail scalars).  With a control block the kernel first closes iteration t - 1.
__global__ void __launch_bounds__(CT) cons_z_kernel(const float* __restrict__ acc, float fN, double pen, int p, float* __restrict__ z, float* __restrict__ out2,
                                                    ConsCtl* __restrict__ c = nullptr, ConsConst k = ConsConst(), double* __restrict__ trace = nullptr,
                                                    int trace_cap = 0)
{
    __shared__ float scratch[33];
    __shared__ int s_stop;
    if (c) {
        if (threadIdx.x == 0) {
            if (!c->stop) {
                const int t = c->t;
                if (c->have_prev) {
                    const double rp_prev = sqrt((double)acc[p + 1]);
                    if (trace && t - 1 < trace_cap) {
                        double* row = trace + 5 * (size_t)(t - 1);
                        row[0] = c->eps_p; row[1] = rp_prev; row[2] = c->eps_d; row[3] = c->rd; row[4] = k.rho;
                        c->trace_rows = t;
                    }
                    if (rp_prev < c->eps_p && c->rd < c->eps_d) { c->niter = t; c->stop = 1; }
                }
                if (!c->stop && t == k.maxit) c->stop = 1;                 // only closing the books: niter stays maxit + 1
                if (!c->stop) {
                    c->eps_p = fmax(sqrt(c->sx2), sqrt(c->sz2) * k.sqrtN) * k.eps_rel + k.sqrt_pN * k.eps_abs;
                    c->eps_d = sqrt((double)acc[p + 2]) * k.eps_rel + k.sqrt_pN * k.eps_abs;
                }
            }
            s_stop = c->stop;
        }
        __syncthreads();
        if (s_stop) return;
    }
    float d2 = 0.f, z2 = 0.f;
    for (int j = threadIdx.x; j < p; j += CT) {
        const float v = acc[j] / fN;
        float zn;
        if ((double)v > pen) zn = (float)((double)v - pen);
        else if ((double)v < -pen) zn = (float)((double)v + pen);
        else zn = 0.f;
        const float dz = zn - z[j];
        d2 += dz * dz; z2 += zn * zn;
        z[j] = zn;
    }
    d2 = block_sum(d2, scratch);
    z2 = block_sum(z2, scratch);
    if (threadIdx.x == 0) {
        out2[0] = d2; out2[1] = z2;
        if (c) {
            c->rd = k.rho * sqrt(k.dN * (double)d2);                       // PADMMLasso.h:151-154
            c->sz2 = (double)z2;
            c->sx2 = (double)acc[p];
            c->have_prev = 1;
            c->t += 1;
        }
    }
}
// r = x - z ; y += frho r ; slots: |r|^2, |y|^2
__global__ void __launch_bounds__(CT) cons_dual_kernel(const float* __restrict__ x, const float* __restrict__ z, float frho, int p, int first,
                                                       float* __restrict__ y, float* __restrict__ slot_r2, float* __restrict__ slot_y2,
                                                       const ConsCtl* __restrict__ c = nullptr)
{
    __shared__ float scratch[33];
    if (c && c->stop) return;
    float r2 = 0.f, y2 = 0.f;
    for (int j = threadIdx.x; j < p; j += CT) {
        const float r = x[j] - z[j];
        const float yn = y[j] + frho * r;
        y[j] = yn;
        r2 += r * r; y2 += yn * yn;
    }
    r2 = block_sum(r2, scratch);
    y2 = block_sum(y2, scratch);
    if (threadIdx.x == 0) {
        *slot_r2 = first ? r2 : *slot_r2 + r2;
        *slot_y2 = first ? y2 : *slot_y2 + y2;
    }
}

struct Block {
    i64 rows = 0, row0 = 0;
    bool tall = true;
    DevBuf<float> Kinv;           // p x ld (tall) or rows x rows (Woodbury)
    DevBuf<float> Ab, x, y;
    DevBuf<float> t1, t2, work;   // Woodbury scratch
};

}  // namespace

void solve_consensus(const LassoRequest& rq, int nthread, b200admm_path* out)
{
    const b200admm_data* d = rq.d;
    Context& c = ctx();
    cudaStream_t s = c.stream;
    Comm& cm = comm();
    if (d->dtype == B200ADMM_F64_DEVICE) throw ArgError("lasso computes in float32: pass f64 host, f32 host or f32 device data");
    const i64 n_local = d->n, p = d->p;
    const int N = nthread;
    if (cm.active() && N != cm.nranks) throw ArgError("nthread must equal the number of ranks of the installed communicator");
    const i64 n = cm.active() ? (i64)std::llround(allreduce_sum_host(s, (double)n_local)) : n_local;
    if (p >= 2147483647LL / 4) throw ArgError("p too large");
    const double t_begin = wall_now();
    const int flag = (rq.standardize ? 1 : 0) + (rq.intercept ? 2 : 0);
    const i64 ld = (p + 3) & ~(i64)3;
    EventTimer tm(s);
    b200admm_timing T;
    memset(&T, 0, sizeof T);

    // ---- ingest + global DataStd (ParLasso.cpp:45-69: standardisation happens before the split) ----
    // One block per rank on device-resident float32 data with both DataStd flags (the multi-GPU deployment): the raw
    // columns are read in place -- statistics, then ONE pass that standardises, forms A'b and writes the fp16 operands
    // of the Gram kernel (gram_std_split_xty, as on the tall path); no standardised float32 copy exists (40 GB per rank
    // at n = 1e6 x p = 8e4 over eight GPUs).
    const char* gram_env = getenv("B200ADMM_GRAM");
    const bool fused_block = cm.active() && d->dtype == B200ADMM_F32_DEVICE && flag == 3 && n_local >= p && !gram_env &&
                             (double)n < 4.0e9 && gram_f16_usable(n_local, p);
    const i64 ldx = (n_local + 3) & ~(i64)3;
    DevBuf<float> Xs, ys(n_local), Xtmp;
    if (!fused_block) Xs.alloc((size_t)ldx * (size_t)p);
    const float* X_in = Xs.p;
    i64 ld_in = ldx;
    tm.start();
    if (!fused_block && ldx != n_local) Xs.zero(s);
    if (d->dtype == B200ADMM_F32_DEVICE) {
        X_in = (const float*)d->x; ld_in = n_local;
        CUDA_CHECK(cudaMemcpyAsync(ys.p, d->y, n_local * sizeof(float), cudaMemcpyDeviceToDevice, s));
    } else {
        float* dst = Xs.p;
        if (ldx != n_local) { Xtmp.alloc((size_t)n_local * (size_t)p); dst = Xtmp.p; X_in = Xtmp.p; ld_in = n_local; }
        ingest_f32(s, d->x, d->dtype, (size_t)n_local * (size_t)p, dst);
        ingest_f32(s, d->y, d->dtype, (size_t)n_local, ys.p);
    }
    T.ingest = tm.stop();
    DevBuf<float> d_meanX(p), d_scaleX(p), d_inv;
    StdStats st;
    tm.start();
    if (fused_block) {
        DevBuf<float> tmp(2 * p + 8);
        d_inv.alloc(p);
        standardize_y_sharded(s, ys.p, n_local, n, flag, st);
        standardize_stats3_sharded(s, X_in, ld_in, n_local, n, p, d_meanX.p, d_scaleX.p, d_inv.p, tmp.p);
        fetch_std_stats(s, p, flag, d_meanX.p, d_scaleX.p, st);
    } else {
        standardize_all(s, X_in, ld_in, Xs.p, ldx, ys.p, n_local, n, p, flag, d_meanX.p, d_scaleX.p, st);
    }
    T.standardize = tm.stop();
    Xtmp.release();

    // ---- row split (PADMMLasso.h:163-178) ---------------------------------------------------------
    std::vector<std::unique_ptr<Block>> blocks;
    if (cm.active()) {
        // rank r must hold exactly the rows the reference gives worker r
        const i64 chunk = n / N;
        const i64 expect = cm.rank < N - 1 ? chunk : chunk + n % N;
        if (n_local != expect) throw ArgError("row-sharded consensus: rank r must hold floor(n/N) rows (the last rank also the remainder)");
        blocks.emplace_back(new Block());
        blocks[0]->rows = n_local; blocks[0]->row0 = 0;
    } else {
        const i64 chunk = n / N;
        if (chunk < 1) throw ArgError("more blocks than observations");
        for (int i = 0; i < N; i++) {
            blocks.emplace_back(new Block());
            blocks[i]->row0 = i * chunk;
            blocks[i]->rows = i < N - 1 ? chunk : chunk + n % N;
        }
    }

    // ---- A_i'b_i, lambda0 = max |X'y| ---------------------------------------------------------------
    DevBuf<float> xy(p);
    DevBuf<unsigned char> Xb;
    tm.start();
    for (size_t i = 0; i < blocks.size(); i++) {
        Block& b = *blocks[i];
        b.tall = b.rows >= p;
        b.Ab.alloc(p); b.x.alloc(p); b.y.alloc(p);
        if (!fused_block) gemv_t<float>(s, Xs.p + b.row0, b.rows, p, ldx, ys.p + b.row0, b.Ab.p);
    }
    if (fused_block) {
        Block& b = *blocks[0];
        Xb.alloc(gram_f16_blocked_bytes(n_local, p));
        DevBuf<float> xty_work(gram_f16_xty_work_floats(n_local, p));
        gram_std_split_xty(s, X_in, n_local, ld_in, p, 0, p, d_meanX.p, d_inv.p, ys.p, Xb.p, b.Ab.p, xty_work.p);
        CUDA_CHECK(cudaMemcpyAsync(xy.p, b.Ab.p, p * sizeof(float), cudaMemcpyDeviceToDevice, s));
    } else {
        gemv_t<float>(s, Xs.p, n_local, p, ldx, ys.p, xy.p);   // X'y over all local rows, as the master does
    }
    allreduce_sum(s, xy.p, p);
    std::vector<float> h_xy(p);
    CUDA_CHECK(cudaMemcpyAsync(h_xy.data(), xy.p, p * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    double lambda0 = 0;
    for (i64 j = 0; j < p; j++) lambda0 = std::max(lambda0, (double)std::fabs(h_xy[j]));

    std::vector<double> lam;
    if (rq.nlambda_given < 1) {
        if (rq.nlambda < 1) throw ArgError("nlambda must be at least 1");
        const double lmax = lambda0 / (double)n * (double)st.scaleY;
        make_lambda_grid(lmax, rq.lmin_ratio > 0 ? rq.lmin_ratio : (n < p ? 0.01 : 1e-4), rq.nlambda, lam);   // (default from the global n)
    } else {
        if (!rq.lambda_given) throw ArgError("lambda is null");
        lam.assign(rq.lambda_given, rq.lambda_given + rq.nlambda_given);
    }
    const int nl = (int)lam.size();
    const double ilam0 = lam[0] * (double)n / (double)st.scaleY;
    const double rho = rq.opts.rho > 0 ? rq.opts.rho : ilam0 / N;           // PADMMLasso.h:199-200
    const float frho = (float)rho;

    // ---- per block: Gram + rho I, explicit inverse (PADMMLasso_Worker::init) --------------------------
    const bool want_tensor = !(gram_env && !strcmp(gram_env, "simt"));
    // (all Gram matrices first, then -- once the standardised copy is no longer needed -- the inverses: at p = 8e4 the
    // copy is 40 GB per rank and the factorisation needs three more p x p arrays next to K_i^-1)
    for (size_t i = 0; i < blocks.size(); i++) {
        Block& b = *blocks[i];
        const float* A = Xs.p + b.row0;
        if (b.tall && fused_block) {
            b.Kinv.alloc((size_t)p * (size_t)ld);
            b.Kinv.zero(s);
            if (!gram_tn_f16_blocked(s, Xb.p, n_local, p, b.Kinv.p, ld)) throw CudaError("fp16 Gram kernel declined the shape");
            if (gram_f16_overflowed(s)) throw CudaError("fp16 Gram split: a standardised value exceeds sqrt(n)");
            add_to_diagonal(s, b.Kinv.p, ld, p, frho);
        } else if (b.tall) {
            b.Kinv.alloc((size_t)p * (size_t)ld);
            b.Kinv.zero(s);
            // (the tensor kernel declines block starts that are not 16-byte aligned; CUDA cores then)
            // flag 3 (centred, unit-norm columns, |x| <= sqrt(n)): fp16 hi / lo operands, otherwise the TF32 split
            const int split = (flag == 3 && (double)n < 4.0e9 && !gram_env) ? GRAM_SPLIT_F16 : GRAM_SPLIT_TF32;
            const bool on_tensor = want_tensor && gram_tn_tensor(s, A, b.rows, ldx, p, b.Kinv.p, ld, split);
            if (on_tensor && split == GRAM_SPLIT_F16 && gram_f16_overflowed(s))
                throw CudaError("fp16 Gram split: a standardised value exceeds sqrt(n)");
            if (!on_tensor)
                gemm<float>(s, true, false, p, p, b.rows, 1.f, A, ldx, A, ldx, 0.f, b.Kinv.p, ld, GEMM_LOWER | GEMM_MIRROR);
            add_to_diagonal(s, b.Kinv.p, ld, p, frho);
        } else {
            const i64 m = b.rows;
            b.Kinv.alloc((size_t)m * (size_t)m);
            gemm<float>(s, false, true, m, m, p, 1.f, A, ldx, A, ldx, 0.f, b.Kinv.p, m, GEMM_LOWER | GEMM_MIRROR);
            add_to_diagonal(s, b.Kinv.p, m, m, frho);
            b.t1.alloc(m); b.t2.alloc(std::max<i64>(m, p)); b.work.alloc(gemv_n_work(m, p));
        }
        b.x.zero(s); b.y.zero(s);
    }
    bool keep_X = false;
    for (auto& b : blocks) if (!b->tall) keep_X = true;
    if (!keep_X) { CUDA_CHECK(cudaStreamSynchronize(s)); Xs.release(); Xb.release(); ys.release(); }
    for (size_t i = 0; i < blocks.size(); i++) {
        Block& b = *blocks[i];
        const i64 m = b.tall ? p : b.rows, ldk = b.tall ? ld : b.rows;
        DevBuf<float> W((size_t)m * (size_t)ldk);
        int info = 0;
        spd_inverse<float>(s, b.Kinv.p, m, ldk, W.p, &info, nullptr);
    }
    T.gram = tm.stop();          // Gram and factorisation of all blocks; reported together

    // ---- iterations ------------------------------------------------------------------------------------
    DevBuf<float> z(p), rhs(p), payload(p + 3), zout2(2), local_tail(2);
    z.zero(s);
    CUDA_CHECK(cudaMemsetAsync(payload.p, 0, (p + 3) * sizeof(float), s));
    CUDA_CHECK(cudaMemsetAsync(local_tail.p, 0, 2 * sizeof(float), s));
    std::vector<float> z_all((size_t)nl * (size_t)p);
    out->niter = (int*)malloc(sizeof(int) * nl);
    out->lambda = (double*)malloc(sizeof(double) * nl);
    if (!out->niter || !out->lambda) throw CodeError(B200ADMM_ENOMEM, "out of host memory");
    TraceRequest& tr = trace_request();
    const double eps_abs = rq.opts.eps_abs, eps_rel = rq.opts.eps_rel;
    const unsigned vg = (unsigned)std::max<i64>(1, std::min<i64>((p + CT - 1) / CT, 64));
    const double sqrt_pN = std::sqrt((double)(p * N));
    const int pi = (int)p;
    // norms of the current iterate: global sum_i |x_i|^2, |z|^2 (sum_i |y_i|^2 arrives with each exchange)
    double sx2 = 0, sz2 = 0;
    // B200ADMM_CONS_BATCH=0: host-driven loop (two device->host reads per iteration) instead of device-controlled batches
    const char* cbatch_env = getenv("B200ADMM_CONS_BATCH");
    const bool batched = !(cbatch_env && !strcmp(cbatch_env, "0"));
    DevBuf<ConsCtl> ctl(1);
    DevBuf<double> trace_dev;

    tm.start();
    for (int k = 0; k < nl; k++) {
        const double lambda = lam[k] * (double)n / (double)st.scaleY;
        const double pen = lambda / (rho * N);
        const bool tracing = tr.buf && tr.cap > 0 && tr.which == k;
        // Pass t performs the x-update of iteration t and one exchange, which also closes the books on
        // iteration t - 1 (its primal residual needed z(t-1), i.e. the previous exchange).
        double eps_p_cur = 0, eps_d_cur = 0, rd_cur = 0;
        bool have_prev = false;
        int niter = rq.opts.maxit + 1;
        if (batched) {
            // passes enqueued in batches without host round trips; the stopping rule runs in cons_z_kernel
            ConsCtl hc;
            memset(&hc, 0, sizeof hc);
            hc.sx2 = sx2; hc.sz2 = sz2; hc.niter = rq.opts.maxit + 1;
            CUDA_CHECK(cudaMemcpyAsync(ctl.p, &hc, sizeof hc, cudaMemcpyHostToDevice, s));
            ConsConst kc;
            kc.eps_abs = eps_abs; kc.eps_rel = eps_rel; kc.sqrt_pN = sqrt_pN; kc.sqrtN = std::sqrt((double)N); kc.rho = rho; kc.dN = (double)N;
            kc.maxit = rq.opts.maxit;
            if (tracing && !trace_dev.p) trace_dev.alloc((size_t)5 * tr.cap);
            for (int t0 = 0; t0 <= rq.opts.maxit && !hc.stop; ) {
                // batch length: passes after the converged one are wasted work (each reads every K_i^-1 once), so a batch
                // is worth about 1.5 ms of device time -- 32 passes at p ~ 1e3, one at p = 8e4 (then still one read, not two)
                const double est_pass = 4.0 * (double)p * (double)p * (double)blocks.size() / 5e12 + 2e-5;
                const int cnt = std::min(std::max(1, std::min(32, (int)(1.5e-3 / est_pass))), rq.opts.maxit + 1 - t0);
                for (int b2 = 0; b2 < cnt; b2++) {
                    for (size_t i = 0; i < blocks.size(); i++) {
                        Block& b = *blocks[i];
                        cons_rhs_kernel<<<vg, CT, 0, s>>>(b.Ab.p, b.y.p, z.p, rho, pi, rhs.p); KERNEL_CHECK();
                        if (b.tall) {
                            gemv_t<float>(s, b.Kinv.p, p, p, ld, rhs.p, b.x.p);
                        } else {
                            const float* A = Xs.p + b.row0;
                            gemv_n<float>(s, A, b.rows, p, ldx, rhs.p, b.t1.p, b.work.p);
                            gemv_t<float>(s, b.Kinv.p, b.rows, b.rows, b.rows, b.t1.p, b.t2.p);
                            gemv_t<float>(s, A, b.rows, p, ldx, b.t2.p, b.x.p);
                            cons_woodbury_kernel<<<vg, CT, 0, s>>>(rhs.p, b.x.p, frho, pi, b.x.p); KERNEL_CHECK();
                        }
                        cons_gather_kernel<<<1, CT, 0, s>>>(b.x.p, b.y.p, frho, pi, i == 0 ? 1 : 0, payload.p, payload.p + p); KERNEL_CHECK();
                    }
                    CUDA_CHECK(cudaMemcpyAsync(payload.p + p + 1, local_tail.p, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
                    allreduce_sum(s, payload.p, (size_t)p + 3);
                    cons_z_kernel<<<1, CT, 0, s>>>(payload.p, (float)N, pen, pi, z.p, zout2.p, ctl.p, kc, tracing ? trace_dev.p : nullptr, tracing ? tr.cap : 0);
                    KERNEL_CHECK();
                    for (size_t i = 0; i < blocks.size(); i++) {
                        Block& b = *blocks[i];
                        cons_dual_kernel<<<1, CT, 0, s>>>(b.x.p, z.p, frho, pi, i == 0 ? 1 : 0, b.y.p, local_tail.p, local_tail.p + 1, ctl.p); KERNEL_CHECK();
                    }
                }
                CUDA_CHECK(cudaMemcpyAsync(&hc, ctl.p, sizeof hc, cudaMemcpyDeviceToHost, s));
                CUDA_CHECK(cudaStreamSynchronize(s));
                t0 += cnt;
            }
            niter = hc.niter;
            sx2 = hc.sx2; sz2 = hc.sz2;
            if (tracing && hc.trace_rows > 0) {
                const int rows = std::min(hc.trace_rows, tr.cap);
                CUDA_CHECK(cudaMemcpy(tr.buf, trace_dev.p, sizeof(double) * 5 * rows, cudaMemcpyDeviceToHost));
                if (tr.nrows) *tr.nrows = rows;
            }
        } else
        for (int t = 0; t <= rq.opts.maxit; t++) {
            for (size_t i = 0; i < blocks.size(); i++) {
                Block& b = *blocks[i];
                cons_rhs_kernel<<<vg, CT, 0, s>>>(b.Ab.p, b.y.p, z.p, rho, pi, rhs.p); KERNEL_CHECK();
                if (b.tall) {
                    gemv_t<float>(s, b.Kinv.p, p, p, ld, rhs.p, b.x.p);                        // K_i^-1 rhs (symmetric)
                } else {
                    const float* A = Xs.p + b.row0;
                    gemv_n<float>(s, A, b.rows, p, ldx, rhs.p, b.t1.p, b.work.p);              // A rhs
                    gemv_t<float>(s, b.Kinv.p, b.rows, b.rows, b.rows, b.t1.p, b.t2.p);        // (AA' + rho I)^-1 (.)
                    gemv_t<float>(s, A, b.rows, p, ldx, b.t2.p, b.x.p);                        // A'(.)
                    cons_woodbury_kernel<<<vg, CT, 0, s>>>(rhs.p, b.x.p, frho, pi, b.x.p); KERNEL_CHECK();
                }
                cons_gather_kernel<<<1, CT, 0, s>>>(b.x.p, b.y.p, frho, pi, i == 0 ? 1 : 0, payload.p, payload.p + p); KERNEL_CHECK();
            }
            // local sum_i |x_i - z|^2 and sum_i |y_i|^2 left by the previous dual update ride along
            CUDA_CHECK(cudaMemcpyAsync(payload.p + p + 1, local_tail.p, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
            allreduce_sum(s, payload.p, (size_t)p + 3);                                         // THE exchange
            float tail[3];
            CUDA_CHECK(cudaMemcpyAsync(tail, payload.p + p, 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaStreamSynchronize(s));
            if (have_prev) {
                const double rp_prev = std::sqrt((double)tail[1]);
                if (tracing && t - 1 < tr.cap) {
                    double* row = tr.buf + 5 * (size_t)(t - 1);
                    row[0] = eps_p_cur; row[1] = rp_prev; row[2] = eps_d_cur; row[3] = rd_cur; row[4] = rho;
                    if (tr.nrows) *tr.nrows = t;
                }
                if (rp_prev < eps_p_cur && rd_cur < eps_d_cur) { niter = t; break; }   // iteration t-1 converged: returns (t-1)+1
            }
            if (t == rq.opts.maxit) break;                                             // only closing the books
            // tolerances of iteration t, from the iterate before it (PADMMBase.h:117-138)
            eps_p_cur = std::max(std::sqrt(sx2), std::sqrt(sz2) * std::sqrt((double)N)) * eps_rel + sqrt_pN * eps_abs;
            eps_d_cur = std::sqrt((double)tail[2]) * eps_rel + sqrt_pN * eps_abs;
            // z-update (identical on every rank), then the dual update of the local blocks
            cons_z_kernel<<<1, CT, 0, s>>>(payload.p, (float)N, pen, pi, z.p, zout2.p); KERNEL_CHECK();
            for (size_t i = 0; i < blocks.size(); i++) {
                Block& b = *blocks[i];
                cons_dual_kernel<<<1, CT, 0, s>>>(b.x.p, z.p, frho, pi, i == 0 ? 1 : 0, b.y.p, local_tail.p, local_tail.p + 1); KERNEL_CHECK();
            }
            float z2[2];
            CUDA_CHECK(cudaMemcpyAsync(z2, zout2.p, 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
            CUDA_CHECK(cudaStreamSynchronize(s));
            rd_cur = rho * std::sqrt((double)N * (double)z2[0]);                        // PADMMLasso.h:151-154
            sz2 = (double)z2[1];
            sx2 = (double)tail[0];
            have_prev = true;
        }
        out->niter[k] = niter;
        out->lambda[k] = lam[k];
        CUDA_CHECK(cudaMemcpyAsync(z_all.data() + (size_t)k * p, z.p, p * sizeof(float), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
    }
    T.iterate = tm.stop();

    tm.start();
    out->nlambda = nl;
    finish_lasso_path(z_all, nl, p, flag, st.meanX, st.scaleX, st.meanY, st.scaleY, out);
    T.finish = tm.stop();
    T.total = wall_now() - t_begin;
    out->rho = rho; out->eig = 0; out->lambda0 = lambda0; out->t = T;
}

}  // namespace b200
