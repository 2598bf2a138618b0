// solver_tall.cu -- host driver of admm_lasso / admm_enet for n > p.
//
// Restates /root/reference/src/Lasso.cpp:39-135 (and Enet.cpp) around the device kernels:
//   ingest (f64 -> f32)            Lasso.cpp:45-50
//   DataStd::standardize           Lasso.cpp:67-68  -> stdize.cu
//   ADMMLassoTall ctor: X'y        ADMMLassoTall.h:164-174 -> gemv_t
//   lambda grid                    Lasso.cpp:78-89
//   init(): Gram, coarse rho, +rho I, LLT      ADMMLassoTall.h:179-216 -> gram / coarse_eig / chol
//   lambda loop with warm starts   Lasso.cpp:97-124 -> ONE launch of the persistent path kernel
//   recover + write_beta_matrix    Lasso.cpp:108-111
//
// With a communicator installed (b200admm_comm_init) `d` is this rank's row block: column sums,
// sums of squares, X'y and the Gram matrix are all-reduced (sum) over NVLink, after which every
// rank holds bit-identical K^-1 and runs the identical iteration -- the serial algorithm of the
// reference, with its O(n p^2) setup split across GPUs.
#include "solvers.h"
#include "kernels.h"
#include "comm.h"
#include <cmath>
#include <cstring>
#include <thread>

namespace b200 {

// DataStd over row shards (a world of one rank degenerates to the plain single-GPU passes).
// y_dev in place; X_in -> X_out (may alias).  Returns meanY / scaleY on the host.
struct StdStats {
    std::vector<float> meanX, scaleX;
    float meanY = 0.f, scaleY = 1.f;
};

// y in place; returns meanY / scaleY on the host (one small synchronising read)
void standardize_y_sharded(cudaStream_t s, float* y, i64 n_local, i64 n_total, int flag, StdStats& st)
{
    if (flag == 0) return;
    DevBuf<float> tmp(4);
    float* ys = tmp.p;          // [0] sum / mean, [1] sumsq / scale
    column_sums<float>(s, y, n_local, 1, n_local, ys);
    allreduce_sum(s, ys, 1);
    mean_from_sums<float>(s, ys, 1, n_total, ys);                       // ys[0] = mean(y)
    column_center_sumsq<float>(s, y, n_local, 1, n_local, ys, ys + 1, false);
    allreduce_sum(s, ys + 1, 1);
    if (flag == 1) {
        scale_from_sumsq<float>(s, ys + 1, 1, n_total, true, ys + 1, nullptr);
        column_apply<float>(s, y, y, n_local, 1, n_local, n_local, nullptr, nullptr, ys + 1);   // y /= scaleY (not centred)
    } else {
        scale_from_sumsq<float>(s, ys + 1, 1, n_total, false, ys + 1, nullptr);
        column_apply<float>(s, y, y, n_local, 1, n_local, n_local, ys, nullptr, ys + 1);        // (y - mean) / scaleY
    }
    float h[2];
    CUDA_CHECK(cudaMemcpyAsync(h, ys, 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    st.meanY = (flag == 1) ? 0.f : h[0];
    st.scaleY = h[1];
}

// a panel of columns: X_in -> X_out (may alias); d_meanX / d_scaleX / tmp point at the panel's entries
// (tmp: 2 * pc floats).  Device work only, no host synchronisation.
static void standardize_cols_sharded(cudaStream_t s, const float* X_in, i64 ld_in, float* X_out, i64 ld_out, i64 n_local, i64 n_total,
                                     i64 pc, int flag, float* d_meanX, float* d_scaleX, float* tmp)
{
    float* sums = tmp;
    float* inv = tmp + pc;
    switch (flag) {
    case 1:
        column_sums<float>(s, X_in, n_local, pc, ld_in, sums);
        allreduce_sum(s, sums, pc);
        mean_from_sums<float>(s, sums, pc, n_total, sums);
        column_center_sumsq<float>(s, const_cast<float*>(X_in), n_local, pc, ld_in, sums, d_scaleX, false);
        allreduce_sum(s, d_scaleX, pc);
        scale_from_sumsq<float>(s, d_scaleX, pc, n_total, true, d_scaleX, inv);
        column_apply<float>(s, X_in, X_out, n_local, pc, ld_in, ld_out, nullptr, inv, nullptr);
        break;
    case 2:
        column_sums<float>(s, X_in, n_local, pc, ld_in, sums);
        allreduce_sum(s, sums, pc);
        mean_from_sums<float>(s, sums, pc, n_total, d_meanX);
        column_apply<float>(s, X_in, X_out, n_local, pc, ld_in, ld_out, d_meanX, nullptr, nullptr);
        break;
    case 3:
        column_sums<float>(s, X_in, n_local, pc, ld_in, sums);
        allreduce_sum(s, sums, pc);
        mean_from_sums<float>(s, sums, pc, n_total, d_meanX);
        column_center_sumsq<float>(s, const_cast<float*>(X_in), n_local, pc, ld_in, d_meanX, d_scaleX, false);
        allreduce_sum(s, d_scaleX, pc);
        scale_from_sumsq<float>(s, d_scaleX, pc, n_total, false, d_scaleX, inv);
        column_apply<float>(s, X_in, X_out, n_local, pc, ld_in, ld_out, d_meanX, inv, nullptr);
        break;
    default:
        if (X_in != X_out) column_apply<float>(s, X_in, X_out, n_local, pc, ld_in, ld_out, nullptr, nullptr, nullptr);
        break;
    }
}

// DataStd flag 3 statistics of a column panel without the apply step (the fp16 Gram path applies them while
// it splits the operands): meanX, scaleX and inv = 1 / scaleX, exactly as in standardize_cols_sharded case 3.
void standardize_stats3_sharded(cudaStream_t s, const float* X_in, i64 ld_in, i64 n_local, i64 n_total, i64 pc,
                                       float* d_meanX, float* d_scaleX, float* d_inv, float* tmp)
{
    float* sums = tmp;
    column_sums<float>(s, X_in, n_local, pc, ld_in, sums);
    allreduce_sum(s, sums, pc);
    mean_from_sums<float>(s, sums, pc, n_total, d_meanX);
    column_center_sumsq<float>(s, const_cast<float*>(X_in), n_local, pc, ld_in, d_meanX, d_scaleX, false);
    allreduce_sum(s, d_scaleX, pc);
    scale_from_sumsq<float>(s, d_scaleX, pc, n_total, false, d_scaleX, d_inv);
}

void fetch_std_stats(cudaStream_t s, i64 p, int flag, const float* d_meanX, const float* d_scaleX, StdStats& st)
{
    st.meanX.assign(p, 0.f);
    st.scaleX.assign(p, 1.f);
    if (flag == 2 || flag == 3) CUDA_CHECK(cudaMemcpyAsync(st.meanX.data(), d_meanX, p * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (flag == 1 || flag == 3) CUDA_CHECK(cudaMemcpyAsync(st.scaleX.data(), d_scaleX, p * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
}

void standardize_all(cudaStream_t s, const float* X_in, i64 ld_in, float* X_out, i64 ld_out, float* y, i64 n_local, i64 n_total, i64 p,
                     int flag, float* d_meanX, float* d_scaleX, StdStats& st)
{
    DevBuf<float> tmp(2 * p + 8);
    standardize_y_sharded(s, y, n_local, n_total, flag, st);
    standardize_cols_sharded(s, X_in, ld_in, X_out, ld_out, n_local, n_total, p, flag, d_meanX, d_scaleX, tmp.p);
    fetch_std_stats(s, p, flag, d_meanX, d_scaleX, st);
}

// column panels of the pipelined host ingest: begin[0] = 0 < begin[1] < ... < begin[npan] = p, every boundary but
// the last a multiple of 256 (whole row blocks of the CTA-pair Gram kernel); pw = width of the wide panels
std::vector<i64> host_panel_schedule(i64 p, i64 pw)
{
    pw = std::max<i64>(256, (pw / 256) * 256);
    std::vector<i64> b;
    for (i64 c = 0; c < p;) {
        b.push_back(c);
        const i64 left = p - c;
        c += (left > pw + 768 || pw <= 256) ? std::min(pw, left) : std::min<i64>(256, left);
    }
    b.push_back(p);
    return b;
}

// DataStd::recover (DataStd.h:183-207) + write_beta_matrix (Lasso.cpp:22-30) for all lambdas.  The lambdas are
// independent: a few host threads rescale them (same arithmetic and order per lambda as recover_sparse) and count their
// non-zeros, then fill their slices of the dgCMatrix arrays (5 ms -> ~1 ms at p = 1e4 x 100 lambdas, which is 5 % of an
// 8-GPU fit).
void finish_lasso_path_inplace(float* c_all, int nl, i64 p, int flag, const std::vector<float>& meanX,
                               const std::vector<float>& scaleX, float meanY, float scaleY, b200admm_path* out);
void finish_lasso_path(const std::vector<float>& z_all, int nl, i64 p, int flag, const std::vector<float>& meanX,
                       const std::vector<float>& scaleX, float meanY, float scaleY, b200admm_path* out)
{
    std::vector<float> c(z_all);
    finish_lasso_path_inplace(c.data(), nl, p, flag, meanX, scaleX, meanY, scaleY, out);
}
// (c_all: nl x p standardised solutions, rescaled in place)
void finish_lasso_path_inplace(float* c_all, int nl, i64 p, int flag, const std::vector<float>& meanX,
                               const std::vector<float>& scaleX, float meanY, float scaleY, b200admm_path* out)
{
    std::vector<float> beta0(nl, 0.f);
    std::vector<size_t> cnt(nl, 0);
    const int nthreads = (int)std::max<i64>(1, std::min<i64>(8, std::min<i64>(nl, ((i64)nl * p) >> 16)));
    auto for_lambdas = [&](auto&& body) {
        if (nthreads <= 1) { for (int k = 0; k < nl; k++) body(k); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++)
            th.emplace_back([&, t]() { for (int k = t; k < nl; k += nthreads) body(k); });
        for (auto& x : th) x.join();
    };
    for_lambdas([&](int k) {
        float* ck = c_all + (size_t)k * p;
        float s = 0.f;
        size_t nz = 0;
        for (i64 j = 0; j < p; j++) {
            if (ck[j] == 0.f) continue;
            if (flag == 1 || flag == 3) ck[j] /= scaleX[j];
            if (flag != 0) ck[j] *= scaleY;
            if (flag == 2 || flag == 3) s += ck[j] * meanX[j];
            if (ck[j] != 0.f) nz++;
        }
        if (flag == 2 || flag == 3) beta0[k] = meanY - s;
        cnt[k] = nz + 1;                                 // the intercept row is always stored
    });
    size_t nnz = 0;
    out->colptr = (int64_t*)malloc(sizeof(int64_t) * (nl + 1));
    if (!out->colptr) throw CodeError(B200ADMM_ENOMEM, "out of host memory");
    for (int k = 0; k < nl; k++) { out->colptr[k] = (int64_t)nnz; nnz += cnt[k]; }
    out->colptr[nl] = (int64_t)nnz;
    out->rowidx = (int*)malloc(sizeof(int) * std::max<size_t>(nnz, 1));
    out->val = (double*)malloc(sizeof(double) * std::max<size_t>(nnz, 1));
    if (!out->rowidx || !out->val) throw CodeError(B200ADMM_ENOMEM, "out of host memory");
    for_lambdas([&](int k) {
        const float* ck = c_all + (size_t)k * p;
        size_t pos = (size_t)out->colptr[k];
        out->rowidx[pos] = 0; out->val[pos] = (double)beta0[k]; pos++;
        for (i64 j = 0; j < p; j++)
            if (ck[j] != 0.f) { out->rowidx[pos] = (int)(j + 1); out->val[pos] = (double)ck[j]; pos++; }
    });
    out->nrow = p + 1;
}

void solve_lasso_like(const LassoRequest& rq, b200admm_path* out)
{
    const b200admm_data* d = rq.d;
    Context& c = ctx();
    cudaStream_t s = c.stream;
    Comm& cm = comm();
    const i64 n_local = d->n, p = d->p;
    const i64 n = cm.active() ? (i64)std::llround(allreduce_sum_host(s, (double)n_local)) : n_local;
    if (!(n > p)) {
        if (cm.active()) throw ArgError("row-sharded lasso needs n > p (the wide solver is not sharded)");
        solve_wide(rq, out);
        return;
    }
    if (d->dtype == B200ADMM_F64_DEVICE) throw ArgError("lasso / enet compute in float32: pass f64 host, f32 host or f32 device data");
    const double t_begin = wall_now();
    const int flag = (rq.standardize ? 1 : 0) + (rq.intercept ? 2 : 0);
    const i64 ld = (p + 3) & ~(i64)3;
    EventTimer tm(s);
    SpanTimer gram_kernel_time(s);
    b200admm_timing T;
    memset(&T, 0, sizeof T);

    // B200ADMM_GRAM: "simt" CUDA-core kernel; "trunc" / "tf32" force a TF32 split.  Default: with DataStd flag 3
    // (centred, unit-norm columns: |x| <= sqrt(n)) the standardised copy is split once into a tile-blocked array
    // of fp16 hi / lo halves and the Gram kernel runs kind::f16 on it (twice the TF32 rate, same 22-bit operands);
    // any other flag keeps the TF32 split, whose 8-bit exponent covers every fp32 column scale.
    const char* gram_env = getenv("B200ADMM_GRAM");
    const bool want_tensor = !(gram_env && !strcmp(gram_env, "simt"));
    const int split_mode = (gram_env && !strcmp(gram_env, "trunc")) ? GRAM_SPLIT_TRUNC : GRAM_SPLIT_TF32;
    const bool use_f16 = want_tensor && !gram_env && flag == 3 && (double)n < 4.0e9 && gram_f16_usable(n_local, p);
    const i64 ldx = (n_local + 3) & ~(i64)3;
    const bool padded = ldx != n_local;
    // fp16 path on device-resident input: the raw columns are read in place (statistics, then one fused
    // standardise + X'y + split pass); no float32 working copy exists at all
    const bool need_copy = !(use_f16 && d->dtype == B200ADMM_F32_DEVICE);
    DevBuf<float> Xs, ys(n_local), Xtmp;
    if (need_copy) Xs.alloc((size_t)ldx * (size_t)p);
    DevBuf<unsigned char> Xb;
    DevBuf<float> d_inv, xty_work;
    if (use_f16) { Xb.alloc(gram_f16_blocked_bytes(n_local, p)); d_inv.alloc(p); }
    DevBuf<float> d_meanX(p), d_scaleX(p);
    DevBuf<float> XY(ld);
    DevBuf<float> G((size_t)p * (size_t)ld);
    StdStats st;
    const bool host_input = d->dtype == B200ADMM_F32_HOST || d->dtype == B200ADMM_F64_HOST;
    const char* pipe_env = getenv("B200ADMM_PIPELINE");
    // (row-sharded runs pipeline too: every rank copies its own rows over its own PCIe link; the per-panel
    // all-reduces of the column statistics need the same panel schedule on every rank, so the panel width is
    // derived from the global row count)
    const bool pipelined = host_input && want_tensor && p >= 1024 && !(pipe_env && !strcmp(pipe_env, "0"));

    if (pipelined) {
        // ---- host input: copy, DataStd, X'y and the Gram matrix pipelined over column panels --------------
        // Columns are standardised independently and row block I of G needs only columns 0 .. 128 (I + 1),
        // so panel k is copied on a second stream while panels < k are standardised and their row blocks
        // of G are computed.  After the last panel lands only its own share of the Gram work remains.
        tm.start();
        if (padded) Xs.zero(s);
        XY.zero(s);
        G.zero(s);
        ingest_f32(s, d->y, d->dtype, (size_t)n_local, ys.p);
        standardize_y_sharded(s, ys.p, n_local, n, flag, st);
        const size_t esz = d->dtype == B200ADMM_F64_HOST ? 8 : 4;
        const i64 n_sched = cm.active() ? (n + cm.nranks - 1) / cm.nranks : n_local;
        i64 pw = (i64)(((size_t)3 << 30) / ((size_t)n_sched * esz));          // ~3 GB of host data per panel
        if (const char* pw_env = getenv("B200ADMM_PANEL_COLS")) pw = atoll(pw_env);     // tests: force several panels
        pw = std::max<i64>(256, (pw / 256) * 256);                        // whole 256-column row blocks of the pair kernel
        // Panel schedule: wide panels first; the last ~1000 columns in single 256-column row blocks.  Whatever
        // Gram work belongs to the final panel starts only after the copy has ended (its share of the whole
        // is 2 w / p for a panel of w columns), and a 256-column launch is no longer wasteful: its <= 40 tiles
        // are cut along K over all CTA pairs.
        const std::vector<i64> pan_begin = host_panel_schedule(p, pw);
        const int npan = (int)pan_begin.size() - 1;
        DevBuf<float> tmp(2 * pw + 8);
        if (use_f16) xty_work.alloc(gram_f16_xty_work_floats(n_local, pw));
        DevBuf<double> slab[2];
        if (esz == 8) { slab[0].alloc((size_t)n_local * (size_t)pw); slab[1].alloc((size_t)n_local * (size_t)pw); }
        // Events and the copy stream are owned by a guard declared AFTER every device buffer the copies land in
        // (Xs, slab[]): on any exit path -- a declined shape, the fp16 overflow flag, a CUDA or NCCL error --
        // its destructor runs first, waits for the copy stream and destroys the events, so no block returns to
        // the cache while a DMA may still be writing into it.
        struct IngestGuard {
            std::vector<cudaEvent_t> landed, freed;
            cudaEvent_t start_ev = nullptr;
            cudaStream_t cs = nullptr;
            ~IngestGuard()
            {
                if (cs) cudaStreamSynchronize(cs);
                for (auto& e : landed) if (e) cudaEventDestroy(e);
                for (auto& e : freed) if (e) cudaEventDestroy(e);
                if (start_ev) cudaEventDestroy(start_ev);
            }
        } ig;
        ig.landed.assign(npan, nullptr);
        ig.freed.assign(2, nullptr);
        ig.cs = copy_stream();
        for (auto& e : ig.landed) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : ig.freed) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&ig.start_ev, cudaEventDisableTiming));
        std::vector<cudaEvent_t>& landed = ig.landed;
        std::vector<cudaEvent_t>& freed = ig.freed;
        cudaStream_t cs = ig.cs;
        CUDA_CHECK(cudaEventRecord(ig.start_ev, s));
        CUDA_CHECK(cudaStreamWaitEvent(cs, ig.start_ev, 0));
        for (int k = 0; k < npan; k++) {
            const i64 c0 = pan_begin[k], pc = pan_begin[k + 1] - c0;
            float* dstp = Xs.p + c0 * ldx;
            if (esz == 4) {
                const float* src = (const float*)d->x + c0 * n_local;
                if (!padded) CUDA_CHECK(cudaMemcpyAsync(dstp, src, (size_t)n_local * pc * 4, cudaMemcpyHostToDevice, cs));
                else CUDA_CHECK(cudaMemcpy2DAsync(dstp, ldx * 4, src, n_local * 4, n_local * 4, pc, cudaMemcpyHostToDevice, cs));
            } else {
                const double* src = (const double*)d->x + c0 * n_local;
                if (k >= 2) CUDA_CHECK(cudaStreamWaitEvent(cs, freed[k & 1], 0));
                CUDA_CHECK(cudaMemcpyAsync(slab[k & 1].p, src, (size_t)n_local * pc * 8, cudaMemcpyHostToDevice, cs));
            }
            CUDA_CHECK(cudaEventRecord(landed[k], cs));
            CUDA_CHECK(cudaStreamWaitEvent(s, landed[k], 0));
            if (esz == 8) {
                if (!padded) convert_f64_to_f32(s, slab[k & 1].p, dstp, (size_t)n_local * pc);
                else for (i64 j = 0; j < pc; j++) convert_f64_to_f32(s, slab[k & 1].p + j * n_local, dstp + j * ldx, (size_t)n_local);
                CUDA_CHECK(cudaEventRecord(freed[k & 1], s));
            }
            bool ok;
            if (use_f16) {
                standardize_stats3_sharded(s, dstp, ldx, n_local, n, pc, d_meanX.p + c0, d_scaleX.p + c0, d_inv.p + c0, tmp.p);
                gram_std_split_xty(s, Xs.p, n_local, ldx, p, c0, c0 + pc, d_meanX.p, d_inv.p, ys.p, Xb.p, XY.p, xty_work.p);
                gram_kernel_time.begin();
                ok = gram_tn_f16_blocked(s, Xb.p, n_local, p, G.p, ld, c0, c0 + pc, k == npan - 1);
            } else {
                standardize_cols_sharded(s, dstp, ldx, dstp, ldx, n_local, n, pc, flag, d_meanX.p + c0, d_scaleX.p + c0, tmp.p);
                gemv_t<float>(s, dstp, n_local, pc, ldx, ys.p, XY.p + c0);
                gram_kernel_time.begin();
                ok = gram_tn_tensor(s, Xs.p, n_local, ldx, p, G.p, ld, split_mode, c0, c0 + pc, k == npan - 1);
            }
            gram_kernel_time.end();
            if (!ok) throw CudaError("pipelined Gram: tensor kernel declined the shape");
        }
        if (use_f16 && gram_f16_overflowed(s)) throw CudaError("fp16 Gram split: a standardised value exceeds sqrt(n)");
        allreduce_sum(s, XY.p, p);
        allreduce_sum(s, G.p, (size_t)p * (size_t)ld);
        fetch_std_stats(s, p, flag, d_meanX.p, d_scaleX.p, st);
        T.gram = tm.stop();                                   // copy + DataStd + X'y + Gram, overlapped
        CUDA_CHECK(cudaStreamSynchronize(cs));
        T.ingest = 0; T.standardize = 0;
    } else {
    // ---- ingest ------------------------------------------------------------------------------
    // The standardised copy keeps a leading dimension that is a multiple of 4 (zero pad rows) so
    // that the TMA descriptor of the tensor-core Gram kernel can address it for any n.
    const float* X_in = Xs.p;
    i64 ld_in = ldx;
    tm.start();
    if (padded && need_copy) Xs.zero(s);
    if (d->dtype == B200ADMM_F32_DEVICE) {
        X_in = (const float*)d->x;                      // standardised out of place, caller's copy untouched
        ld_in = n_local;
        CUDA_CHECK(cudaMemcpyAsync(ys.p, d->y, n_local * sizeof(float), cudaMemcpyDeviceToDevice, s));
    } else {
        float* dst = Xs.p;
        if (padded) { Xtmp.alloc((size_t)n_local * (size_t)p); dst = Xtmp.p; X_in = Xtmp.p; ld_in = n_local; }
        ingest_f32(s, d->x, d->dtype, (size_t)n_local * (size_t)p, dst);
        ingest_f32(s, d->y, d->dtype, (size_t)n_local, ys.p);
    }
    T.ingest = tm.stop();

    // ---- DataStd -----------------------------------------------------------------------------
    tm.start();
    if (use_f16) {
        // statistics only: the apply step is fused with X'y and the operand split below
        DevBuf<float> tmp(2 * p + 8);
        standardize_y_sharded(s, ys.p, n_local, n, flag, st);
        standardize_stats3_sharded(s, X_in, ld_in, n_local, n, p, d_meanX.p, d_scaleX.p, d_inv.p, tmp.p);
        fetch_std_stats(s, p, flag, d_meanX.p, d_scaleX.p, st);
    } else {
        standardize_all(s, X_in, ld_in, Xs.p, ldx, ys.p, n_local, n, p, flag, d_meanX.p, d_scaleX.p, st);
    }
    T.standardize = tm.stop();
    if (!use_f16) Xtmp.release();

    // ---- X'y, lambda0, Gram ------------------------------------------------------------------
    tm.start();
    XY.zero(s);
    if (!use_f16) gemv_t<float>(s, Xs.p, n_local, p, ldx, ys.p, XY.p);
    G.zero(s);
    bool on_tensor = false;
    if (use_f16) {
        xty_work.alloc(gram_f16_xty_work_floats(n_local, p));
        gram_std_split_xty(s, X_in, n_local, ld_in, p, 0, p, d_meanX.p, d_inv.p, ys.p, Xb.p, XY.p, xty_work.p);
        Xtmp.release();
        allreduce_sum(s, XY.p, p);
        gram_kernel_time.begin();
        on_tensor = gram_tn_f16_blocked(s, Xb.p, n_local, p, G.p, ld);
        gram_kernel_time.end();
        if (!on_tensor) throw CudaError("fp16 Gram kernel declined the shape");
        if (gram_f16_overflowed(s)) throw CudaError("fp16 Gram split: a standardised value exceeds sqrt(n)");
    } else {
        allreduce_sum(s, XY.p, p);
        gram_kernel_time.begin();
        on_tensor = want_tensor && gram_tn_tensor(s, Xs.p, n_local, ldx, p, G.p, ld, split_mode);
        if (!on_tensor)     // CUDA-core path (shapes the tensor kernel does not take)
            gemm<float>(s, true, false, p, p, n_local, 1.f, Xs.p, ldx, Xs.p, ldx, 0.f, G.p, ld, GEMM_LOWER | GEMM_MIRROR);
        gram_kernel_time.end();
    }
    allreduce_sum(s, G.p, (size_t)p * (size_t)ld);
    T.gram = tm.stop();
    }
    std::vector<float> h_xy(p);
    {
        CaptureRequest& cap = capture_request();
        if (cap.gram) CUDA_CHECK(cudaMemcpy2DAsync(cap.gram, (size_t)p * 4, G.p, (size_t)ld * 4, (size_t)p * 4, (size_t)p, cudaMemcpyDeviceToHost, s));
        if (cap.xy) CUDA_CHECK(cudaMemcpyAsync(cap.xy, XY.p, p * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (cap.stats) {
            std::copy(st.meanX.begin(), st.meanX.end(), cap.stats);
            std::copy(st.scaleX.begin(), st.scaleX.end(), cap.stats + p);
            cap.stats[2 * p] = st.meanY; cap.stats[2 * p + 1] = st.scaleY;
        }
    }
    CUDA_CHECK(cudaMemcpyAsync(h_xy.data(), XY.p, p * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    g_last_gram_seconds = gram_kernel_time.total();
    Xs.release();                                       // the tall solver never touches X again
    Xb.release();
    ys.release();

    float lambda0 = 0.f;
    for (i64 j = 0; j < p; j++) lambda0 = std::max(lambda0, std::fabs(h_xy[j]));
    if (rq.enet) lambda0 = (float)((double)lambda0 / ((double)(float)rq.alpha + 0.0001));   // ADMMEnet.h:56

    // ---- lambda sequence (Lasso.cpp:78-89) -----------------------------------------------------
    std::vector<double> lam;
    if (rq.nlambda_given < 1) {
        if (rq.nlambda < 1) throw ArgError("nlambda must be at least 1");
        const double lmax = (double)lambda0 / (double)n * (double)st.scaleY;
        // lmin_ratio <= 0: the front end's default (R/30_admm_lasso.R: 0.01 if n < p else 1e-4) taken from the GLOBAL row count --
        // a rank of a row-sharded run may hold fewer rows than columns
        make_lambda_grid(lmax, rq.lmin_ratio > 0 ? rq.lmin_ratio : (n < p ? 0.01 : 1e-4), rq.nlambda, lam);
    } else {
        if (!rq.lambda_given) throw ArgError("lambda is null");
        lam.assign(rq.lambda_given, rq.lambda_given + rq.nlambda_given);
    }
    const int nl = (int)lam.size();
    std::vector<double> ilam(nl);
    for (int k = 0; k < nl; k++) ilam[k] = lam[k] * (double)n / (double)st.scaleY;           // Lasso.cpp:99

    // ---- rho (ADMMLassoTall.h:194-202) ---------------------------------------------------------
    double rho = rq.opts.rho;
    double ev = 0.0;
    tm.start();
    if (rho <= 0) {
        const float evf = coarse_eig_device(s, G.p, p, ld, nullptr);
        ev = (double)evf;
        const float lambda_f = (float)ilam[0];
        rho = std::pow((double)evf, 1.0 / 3) * std::pow((double)lambda_f, 2.0 / 3);
    }
    T.eig = tm.stop();

    // ---- K^-1 = (X'X + rho I)^-1  (replaces LLT::compute, ADMMLassoTall.h:204-205) --------------
    tm.start();
    add_to_diagonal(s, G.p, ld, p, (float)rho);
    {
        // (+64: a size class of its own, so the block cache can never hand G's block out as W or the reverse --
        // the cached factor graph is keyed on both pointers)
        DevBuf<float> W((size_t)p * (size_t)ld + 64);
        int info = 0;
        spd_inverse_f32(s, G.p, p, ld, W.p, &info);
    }
    T.factor = tm.stop();

    // ---- the lambda path: one persistent kernel ------------------------------------------------
    // Row-sharded runs: every rank holds the same K^-1, so the iterations are sharded too -- rank r streams
    // rows [p r / N, p (r + 1) / N) only and the per-iteration exchange (new z / y entries, partial norms, a
    // barrier flag) goes through cudaIpc-mapped peer memory inside the kernel (fadmm_tall.cu).  State and
    // z_out then live in the exported block.  B200ADMM_SHARD_ITER=0 (or a failed mapping) keeps the replicated
    // iterations.
    const char* shard_env = getenv("B200ADMM_SHARD_ITER");
    const size_t state_floats = tall_state_floats((int)p);
    const size_t off_flags = (state_floats + 3) & ~(size_t)3, off_zout = off_flags + 32;
    PeerBlock* pb = nullptr;
    if (cm.active() && cm.nranks <= 8 && !(shard_env && !strcmp(shard_env, "0"))) {
        PeerBlock& blk = peer_block(s, off_zout + (size_t)nl * (size_t)p);
        if (blk.ok) pb = &blk;
    }
    DevBuf<float> state, z_out;
    DevBuf<int> abort_dev(1);
    if (!pb) { state.alloc(state_floats); z_out.alloc((size_t)nl * (size_t)p); }
    float* const state_p = pb ? pb->local : state.p;
    float* const z_out_p = pb ? pb->local + off_zout : z_out.p;
    DevBuf<int> niter_dev(nl), trace_rows(1);
    DevBuf<double> lam_dev(nl);
    DevBuf<unsigned long long> barrier(1);
    TraceRequest& tr = trace_request();
    DevBuf<double> trace_dev;
    const bool tracing = tr.buf && tr.cap > 0 && tr.which >= 0 && tr.which < nl;
    if (tracing) trace_dev.alloc((size_t)5 * tr.cap);

    if (pb) {
        // nobody may store into a peer's block before that peer has cleared it
        CUDA_CHECK(cudaMemsetAsync(pb->local, 0, off_zout * sizeof(float), s));
        abort_dev.zero(s);
        CUDA_CHECK(cudaStreamSynchronize(s));
        (void)allreduce_sum_host(s, 0.0);
    } else {
        state.zero(s);
    }
    barrier.zero(s);
    trace_rows.zero(s);
    CUDA_CHECK(cudaMemcpyAsync(lam_dev.p, ilam.data(), nl * sizeof(double), cudaMemcpyHostToDevice, s));

    TallPathArgs a;
    a.Kinv = G.p; a.XY = XY.p; a.lambdas = lam_dev.p; a.nl = nl; a.p = (int)p; a.maxit = rq.opts.maxit;
    a.eps_abs = rq.opts.eps_abs; a.eps_rel = rq.opts.eps_rel; a.rho = rho;
    a.enet = rq.enet ? 1 : 0; a.alpha = rq.alpha;
    a.state = state_p; a.z_out = z_out_p; a.niter_out = niter_dev.p;
    if (pb) {
        a.nranks = cm.nranks; a.rank = cm.rank;
        for (int k = 0; k < cm.nranks; k++) a.peers[k] = pb->peers[k];
        a.off_flags = off_flags; a.off_zout = off_zout; a.abort_flag = abort_dev.p;
    }
    a.trace = tracing ? trace_dev.p : nullptr; a.trace_cap = tracing ? tr.cap : 0; a.trace_lambda = tracing ? tr.which : -1;
    a.trace_rows = trace_rows.p; a.barrier = barrier.p;
    DevBuf<unsigned long long> prof_dev;
    if (getenv("B200ADMM_PATH_PROF")) { prof_dev.alloc(8); prof_dev.zero(s); a.prof = prof_dev.p; }
    const char* snake_env = getenv("B200ADMM_SNAKE");
    a.snake = snake_env ? atoi(snake_env) : 1;
    // one GPU: read one triangle of the symmetric K^-1 per iteration (B200ADMM_TALL_TRI=0: the full-row kernel)
    DevBuf<float> tri_part;
    const char* tri_env = getenv("B200ADMM_TALL_TRI");
    if (!pb && !(tri_env && !strcmp(tri_env, "0"))) {
        const size_t nf = tall_tri_part_floats((int)p);
        if (nf) { tri_part.alloc(nf); a.tri_part = tri_part.p; }
    }
    tm.start();
    launch_tall_path(s, a);
    T.iterate = tm.stop();

    // ---- results --------------------------------------------------------------------------------
    tm.start();
    // (pinned staging buffer kept between calls: the 4 MB of solutions come back at PCIe speed instead of through the
    // driver's pageable path, and are rescaled where they land)
    static float* z_host = nullptr;
    static size_t z_host_floats = 0;
    const size_t z_floats = (size_t)nl * (size_t)p;
    if (z_floats > z_host_floats) {
        if (z_host) cudaFreeHost(z_host);
        z_host = nullptr; z_host_floats = 0;
        CUDA_CHECK(cudaHostAlloc((void**)&z_host, z_floats * sizeof(float), cudaHostAllocDefault));
        z_host_floats = z_floats;
    }
    out->niter = (int*)malloc(sizeof(int) * nl);
    out->lambda = (double*)malloc(sizeof(double) * nl);
    if (!out->niter || !out->lambda) throw CodeError(B200ADMM_ENOMEM, "out of host memory");
    CUDA_CHECK(cudaMemcpyAsync(z_host, z_out_p, z_floats * sizeof(float), cudaMemcpyDeviceToHost, s));
    int aborted = 0;
    if (pb) CUDA_CHECK(cudaMemcpyAsync(&aborted, abort_dev.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaMemcpyAsync(out->niter, niter_dev.p, nl * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (tracing) {
        int rows = 0;
        CUDA_CHECK(cudaMemcpyAsync(&rows, trace_rows.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        rows = std::min(rows, tr.cap);
        CUDA_CHECK(cudaMemcpyAsync(tr.buf, trace_dev.p, (size_t)5 * rows * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (tr.nrows) *tr.nrows = rows;
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
    if (prof_dev.p) {
        unsigned long long h[8];
        CUDA_CHECK(cudaMemcpy(h, prof_dev.p, sizeof h, cudaMemcpyDeviceToHost));
        const double it = (double)std::max<unsigned long long>(1, h[7]);
        fprintf(stderr, "[b200admm path profile, rank %d, CTA 0] cycles / iteration: A %.0f  B %.0f  sync+fence %.0f  local barrier %.0f  "
                        "peer wait %.0f  C scalars %.0f  C total %.0f  (%llu iterations)\n", cm.rank, h[0] / it, h[1] / it, h[2] / it, h[3] / it,
                h[4] / it, h[5] / it, h[6] / it, h[7]);
    }
    if (aborted) throw CodeError(B200ADMM_ENCCL, "sharded lambda path: a peer rank did not reach the iteration barrier in time");
    for (int k = 0; k < nl; k++) out->lambda[k] = lam[k];
    out->nlambda = nl;
    finish_lasso_path_inplace(z_host, nl, p, flag, st.meanX, st.scaleX, st.meanY, st.scaleY, out);
    T.finish = tm.stop();
    T.total = wall_now() - t_begin;
    out->rho = rho; out->eig = ev; out->lambda0 = lambda0; out->t = T;
}

}  // namespace b200
