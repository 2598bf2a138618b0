// kernels.h -- host-callable launchers of the CUDA kernels in this directory.
#pragma once
#include "common.cuh"

namespace b200 {

// ---- gemm.cu -----------------------------------------------------------------------------
enum GemmMode {
    GEMM_LOWER        = 1,    // compute / store only C(i,j) with i >= j (square C)
    GEMM_MIRROR       = 2,    // additionally store C(j,i) = C(i,j)  (with GEMM_LOWER)
    GEMM_A_LOWER_TRI  = 4,    // op(A)=A,  A lower triangular: skip + mask k > i
    GEMM_AT_LOWER_TRI = 8,    // op(A)=A', A lower triangular: skip + mask k < i
    GEMM_B_LOWER_TRI  = 16,   // op(B)=B,  B lower triangular: skip + mask k < j
    GEMM_BT_LOWER_TRI = 32,   // op(B)=B', B lower triangular: skip + mask k > j
};
template <class T>
void gemm(cudaStream_t s, bool ta, bool tb, i64 M, i64 N, i64 K, T alpha, const T* A, i64 lda,
          const T* B, i64 ldb, T beta, T* C, i64 ldc, int mode);

// ---- gram_tc.cu : X'X on the tcgen05 tensor cores (split products, fp32-accurate) ----------
// G (p x p, leading dimension ld) must be zeroed by the caller; the full symmetric matrix is written.
// split selects how every fp32 operand is cut into two tensor-core operands (three products per k):
enum GramSplit {
    GRAM_SPLIT_TRUNC = 0,   // TF32, hi = the raw word (the tensor core ignores the 13 low mantissa bits)
    GRAM_SPLIT_TF32  = 1,   // TF32, hi / lo rounded to nearest and rewritten in shared memory (any fp32 data)
    GRAM_SPLIT_F16   = 2,   // fp16 hi / lo (kind::f16, twice the TF32 rate): data of unit scale only, i.e. columns
                            // standardised by DataStd; out-of-range values raise a flag -> gram_f16_overflowed()
};
// returns false if the shape cannot use the tensor path (caller falls back to gemm<float>)
// [col_begin, col_end): only the row blocks of G belonging to these columns of X are computed (they need
// columns 0 .. col_end of X); mirror = false defers the lower -> upper copy to the last panel.
bool gram_tn_tensor(cudaStream_t s, const float* X, i64 n, i64 ldx, i64 p, float* G, i64 ld, int split,
                    i64 col_begin = 0, i64 col_end = -1, bool mirror = true);
// fp16 path in two steps, for callers that build the Gram matrix panel by panel (the tall solver):
//   gram_split_f16_blocked : columns [col_begin, col_end) of X (n x p, column-major, ldx % 4 == 0) -> the blocked
//                            fp16 hi | lo operand array Xb (gram_f16_blocked_bytes(n, p) bytes; 128-column x
//                            32-row boxes of 16 KB, one TMA burst each); col_begin % 128 == 0
//   gram_tn_f16_blocked    : same contract as gram_tn_tensor, operands read from Xb (columns < col_end split)
bool gram_f16_usable(i64 n, i64 p);
void gram_f16_plan(int ntiles, int npairs, int nk, int* cover, long long* per_pair, int* nslices, int* split_tiles);
size_t gram_f16_blocked_bytes(i64 n, i64 p);
void gram_split_f16_blocked(cudaStream_t s, const float* X, i64 n, i64 ldx, i64 p, i64 col_begin, i64 col_end, void* Xb);
// DataStd flag 3 fused with the split: reads the RAW columns [col_begin, col_end) of X once, applies
// (x - mean[j]) * inv[j] (col_apply's arithmetic), writes the blocked operands and xty[j] = sum_i x_std(i,j) y(i)
// (y: the standardised response).  work: gram_f16_xty_work_floats(n, col_end - col_begin) floats.
size_t gram_f16_xty_work_floats(i64 n, i64 ncols);
void gram_std_split_xty(cudaStream_t s, const float* X, i64 n, i64 ldx, i64 p, i64 col_begin, i64 col_end,
                        const float* mean, const float* inv, const float* y, void* Xb, float* xty, float* work);
bool gram_tn_f16_blocked(cudaStream_t s, const void* Xb, i64 n, i64 p, float* G, i64 ld,
                         i64 col_begin = 0, i64 col_end = -1, bool mirror = true);
// C (p x p, ld) -= X'X on the lower-triangle tiles (3xTF32 CTA-pair kernel, subtract epilogue; the strict upper
// part of the diagonal tiles is overwritten with values nobody reads): the rank-n update of the blocked Cholesky.
// X: n x p column-major, ldx % 4 == 0, p >= 256.  Returns false if the shape is not taken.
bool gram_tn_tensor_sub(cudaStream_t s, const float* X, i64 n, i64 ldx, i64 p, float* C, i64 ld);
// General two-operand form on the same CTA-pair 3xTF32 kernel:  C (M x N, ldc) op= A' B,  A: K x M, B: K x N, both
// column-major (lda, ldb multiples of 4, 16-byte aligned) -- the TN product whose operands are K-major as they are.
//   tile_mode  TN_TILES_ALL: every 256 x 256 tile; TN_TILES_LOWER: tiles with J <= I; TN_TILES_UPPER: J >= I (M == N)
//   klo_mode / khi_mode: the K range of tile (I, J) in 256-row blocks -- first block 0 / I / J / max(I, J), one past
//              the last block = all / I + 1 / J + 1 / min(I, J) + 1 -- so that the zero blocks of a triangular operand
//              are skipped (its diagonal blocks must hold explicit zeros)
//   epi        TN_STORE: C = A'B; TN_SUB: C -= A'B; TN_NEG: C = -A'B
// Returns false if the shape / device cannot take it.
enum { TN_TILES_ALL = 0, TN_TILES_LOWER = 1, TN_TILES_UPPER = 2 };
enum { TN_K_ALL = 0, TN_K_I = 1, TN_K_J = 2, TN_K_OUTER = 3 };     // klo: OUTER = max(I, J); khi: OUTER = min(I, J)
enum { TN_STORE = 0, TN_SUB = 1, TN_NEG = 2 };
bool gemm_tn_tensor(cudaStream_t s, const float* A, i64 lda, const float* B, i64 ldb, i64 M, i64 N, i64 K, float* C, i64 ldc,
                    int tile_mode, int klo_mode, int khi_mode, int epi);
void mirror_lower_to_upper(cudaStream_t s, float* G, i64 p, i64 ld);
// synchronises `s`; true (and the flag is cleared) if an fp16 split since the last call met |x| > 65000
bool gram_f16_overflowed(cudaStream_t s);

// ---- gemv.cu -----------------------------------------------------------------------------
// out[j] = sum_i A(i,j) v[i]   (A m x ncol column-major, lda)  -- one dot product per column
template <class T> void gemv_t(cudaStream_t s, const T* A, i64 m, i64 ncol, i64 lda, const T* v, T* out);
// out[list[k]] = A(:, list[k])' v for k < count, bit-identical to gemv_t's value when gemv_t_uses_warp_kernel(m)
bool gemv_t_uses_warp_kernel(i64 m);
template <class T> void gemv_t_list(cudaStream_t s, const T* A, i64 m, i64 lda, const T* v, const int* list, int count, T* out);
// out[i] = sum_j A(i,j) v[j]   -- row-parallel; `work` must hold gemv_n_work(m, ncol) entries
template <class T> void gemv_n(cudaStream_t s, const T* A, i64 m, i64 ncol, i64 lda, const T* v, T* out, T* work);
size_t gemv_n_work(i64 m, i64 ncol);

// ---- stdize.cu ---------------------------------------------------------------------------
// DataStd (/root/reference/src/DataStd.h:89-155).  flag = standardize + 2*intercept.
// X_out may alias X_in.  meanX / scaleX: device arrays of length p (written as the flag requires).
// tmp: 2 * p scratch entries (device).
template <class T>
void standardize_columns(cudaStream_t s, const T* X_in, T* X_out, i64 n, i64 p, i64 ld, int flag,
                         T* meanX, T* scaleX, T* tmp);
// y: device vector, in place.  out2 (device): {meanY, scaleY}; tmp: 2 scratch entries
template <class T> void standardize_y(cudaStream_t s, T* y, i64 n, int flag, T* out2, T* tmp);
// column sums / sums of squared deviations for row-sharded standardisation
template <class T> void column_sums(cudaStream_t s, const T* X, i64 n, i64 p, i64 ld, T* sums);
template <class T> void column_center_sumsq(cudaStream_t s, T* X, i64 n, i64 p, i64 ld, const T* mean, T* sumsq, bool center_in_place);
template <class T> void column_scale(cudaStream_t s, T* X, i64 n, i64 p, i64 ld, const T* inv_scale);
void convert_f64_to_f32(cudaStream_t s, const double* in, float* out, size_t count);

// ---- chol.cu -----------------------------------------------------------------------------
// In-place lower Cholesky of the p x p matrix A (column-major, lda); `work` >= chol_work(p).
// Also leaves the inverses of the diagonal blocks in `work` for spd_inverse.
// info (device int): 0 ok, k > 0: pivot k not positive.
template <class T> size_t chol_work(i64 p);
template <class T> void chol_lower(cudaStream_t s, T* A, i64 p, i64 lda, T* work, int* info_dev);
// W (p x p, zero-initialised by the callee) <- L^-1 given the factor in A and chol's `work`
size_t tri_inverse_tmp(i64 p);      // entries of tri_inverse_lower's `tmp`
template <class T> void tri_inverse_lower(cudaStream_t s, const T* L, i64 p, i64 lda, const T* work, T* W, i64 ldw, T* tmp);
// float32, p >= 512: a <- a^-1 (full symmetric) with all O(p^3) work on the tcgen05 3xTF32 kernel (upper-form blocked
// Cholesky, block inverse with K-range skipping, Z'Z); Z (p x ld) is left holding L^-1, Zt (p x ld) is scratch.
// Everything is enqueued on `s` (capturable into a CUDA graph); info_dev as in chol_lower.
size_t spd_tc_work_floats(i64 p);
bool spd_inverse_tc_usable(i64 p, i64 ld);
void spd_inverse_tc(cudaStream_t s, float* a, i64 p, i64 ld, float* Z, float* Zt, float* work, int* info_dev);
double diag_kernel_bench(cudaStream_t s, int mode, int reps, float* A, i64 lda, float* Dinv, int* info);
// Kinv (p x p, full symmetric) <- W' W
template <class T> void gram_of_lower(cudaStream_t s, const T* W, i64 p, i64 ldw, T* Kinv, i64 ldk);
// solve L L' x = b in place for one right-hand side (LAD's get_x; not on the per-iteration path)
template <class T> void chol_solve_vec(cudaStream_t s, const T* L, i64 p, i64 lda, T* b);

// ---- fadmm_tall.cu : the persistent lambda-path kernel (lasso / enet, n > p) ---------------
struct TallPathArgs {
    const float* Kinv;      // p x p, full symmetric inverse of X'X + rho I
    const float* XY;        // p
    const double* lambdas;  // nl, internal scale (lambda * n / scaleY)
    int nl;
    int p;
    int maxit;
    double eps_abs, eps_rel, rho;
    int enet;
    double alpha;
    float* state;           // workspace, tall_state_floats(p) floats, zeroed by the caller for a cold start
    float* z_out;           // nl x p  (row k = z at lambda k, standardised scale)
    int* niter_out;         // nl
    double* trace;          // optional: rows of 5 doubles for lambda `trace_lambda`
    int trace_cap;
    int trace_lambda;
    int* trace_rows;        // device int
    unsigned long long* barrier;   // device counter, zeroed
    int snake;              // alternate the row sweep direction every iteration (L2 reuse)
    // row-sharded runs: state / z_out live in one cudaIpc-exported block per rank, mapped into every peer
    int nranks = 1, rank = 0;
    float* peers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // block base of each rank (peers[rank] == state)
    size_t off_flags = 0;   // float offset of the 8 x u64 barrier flags inside a block
    size_t off_zout = 0;    // float offset of z_out inside a block
    int* abort_flag = nullptr;   // device int (local), zeroed; set when a peer did not arrive in time
    unsigned long long* prof = nullptr;   // optional (B200ADMM_PATH_PROF=1): 8 cycle counters of CTA 0, summed over the path
    // single GPU: workspace of tall_tri_part_floats(p) floats selects the kernel that reads one triangle of Kinv per
    // iteration (nullptr: the full-row kernel)
    float* tri_part = nullptr;
};
size_t tall_state_floats(int p);
size_t tall_tri_part_floats(int p);   // 0: p too large for the one-triangle kernel
int tall_tri_plan(int p, int sms, int* rows, int cap, long long* smem_bytes);   // host only: grid size (0: not taken, -1: cap)
// returns the grid size used
int launch_tall_path(cudaStream_t s, const TallPathArgs& a);

// stand-alone fused z+u+residual pass (the K8 micro-benchmark of SURVEY.md section 8d)
void fused_zu_pass(cudaStream_t s, const float* x, const float* adj_y, const float* old_z, const float* adj_z,
                   float* z, float* y, i64 len, double lambda, double rho, int enet, double alpha,
                   double* sums_dev /* 6 */, float* partial_work /* >= 6 * fused_zu_blocks() */);
int fused_zu_blocks();

// ---- synth.cu ----------------------------------------------------------------------------
void synth_design_f32(cudaStream_t s, float* X, float* y, i64 nrows, i64 p, i64 row0, uint64_t seed,
                      float mean_x, float sd_x, int nsig, float noise);

}  // namespace b200
