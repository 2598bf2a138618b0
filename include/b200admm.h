/* b200admm.h -- C ABI of libb200admm.so, the B200 (sm_100a) drop-in for the native entry
 * points of the R package yixuan/ADMM.
 *
 * Each function below replaces one `RcppExport SEXP` entry of the reference; the Rcpp glue a
 * maintainer would add (SEXP <-> these plain structs) is shown in INTEGRATION.md.  Citations
 * are into /root/reference.
 *
 *   b200admm_lasso     <- admm_lasso     src/Lasso.cpp:32-138     (R: R/30_admm_lasso.R:139-146)
 *   b200admm_enet      <- admm_enet      src/Enet.cpp:31-138      (R: R/40_admm_enet.R:53-62)
 *   b200admm_parlasso  <- admm_parlasso  src/ParLasso.cpp:33-111  (R: R/30_admm_lasso.R:148-157)
 *   b200admm_lad       <- admm_lad       src/LAD.cpp:16-48        (R: R/20_admm_lad.R:60-67)
 *   b200admm_bp        <- admm_bp        src/BP.cpp:20-46         (R: R/10_admm_bp.R:103-118)
 *
 * Conventions: plain pointers and sizes only; matrices are column-major (R layout); every
 * function returns 0 on success or a negative B200ADMM_E* code, and never throws across the
 * boundary (the reference's BEGIN_RCPP/END_RCPP turns C++ exceptions into R errors; the glue
 * does `if (rc) Rcpp::stop(b200admm_last_error())`).  Result objects are allocated by the
 * library and released with the matching b200admm_free_*.  There is no CPU fallback: if no
 * CUDA device is usable the call fails with B200ADMM_ENODEVICE.
 */
#ifndef B200ADMM_H
#define B200ADMM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ADMM_VERSION 100

/* error codes */
#define B200ADMM_OK          0
#define B200ADMM_EINVAL     -1   /* bad argument (the R front-ends' stop() conditions) */
#define B200ADMM_ENODEVICE  -2   /* no usable CUDA device / wrong architecture */
#define B200ADMM_ECUDA      -3   /* CUDA runtime error, text in b200admm_last_error() */
#define B200ADMM_ENOTSPD    -4   /* Cholesky met a non-positive pivot */
#define B200ADMM_ELANCZOS   -5   /* fewer than 3 rows/cols: Spectra's ctor throws (SymEigsSolver.h:478-482) */
#define B200ADMM_ENOMEM     -6
#define B200ADMM_ENCCL      -7

/* where the caller's x / y live */
#define B200ADMM_F64_HOST    0   /* R's REALSXP: double, host, column-major (what .Call receives) */
#define B200ADMM_F32_HOST    1   /* float, host (pinned or pageable) */
#define B200ADMM_F32_DEVICE  2   /* float, already in HBM on the current device; never modified */
#define B200ADMM_F64_DEVICE  3   /* double, in HBM (LAD / BP) */

typedef struct b200admm_data {
    int64_t     n;        /* rows (observations) */
    int64_t     p;        /* columns (variables) */
    int         dtype;    /* B200ADMM_F64_HOST ... */
    const void* x;        /* n x p, column-major, leading dimension n */
    const void* y;        /* n */
} b200admm_data;

/* the R-side `opts` list (R/30_admm_lasso.R:143-146; src/Lasso.cpp:59-63) */
typedef struct b200admm_opts {
    int    maxit;      /* default 10000 */
    double eps_abs;    /* 1e-5 (lasso/enet), 1e-4 (lad/bp) */
    double eps_rel;
    double rho;        /* <= 0: automatic (lasso/enet); lad/bp default 1 */
} b200admm_opts;

/* phase timings in seconds, measured with CUDA events on the library's stream */
typedef struct b200admm_timing {
    double ingest;       /* host->device copy + f64->f32 conversion (0 for device input) */
    double standardize;  /* DataStd                                                   */
    double gram;         /* X'y, X'X (or XX')                                         */
    double eig;          /* coarse Lanczos                                            */
    double factor;       /* Cholesky (+ explicit inverse / M = L^-1 A)                */
    double iterate;      /* the ADMM loop(s), all lambdas                             */
    double finish;       /* device->host of the solutions, recover, CSC assembly      */
    double total;        /* wall clock of the whole call                              */
} b200admm_timing;

/* result of lasso / enet / parlasso: List(lambda, beta = dgCMatrix (p+1) x nlambda, niter)
 * (src/Lasso.cpp:131-135).  Row 0 of beta is the intercept and is always stored. */
typedef struct b200admm_path {
    int      nlambda;
    int64_t  nrow;       /* p + 1 */
    double*  lambda;     /* [nlambda] */
    int*     niter;      /* [nlambda]; maxit + 1 when the loop ran out (src/FADMMBase.h:264) */
    int64_t* colptr;     /* [nlambda + 1]   -- dgCMatrix @p */
    int*     rowidx;     /* [nnz]           -- dgCMatrix @i */
    double*  val;        /* [nnz]           -- dgCMatrix @x */
    double   rho;        /* rho actually used (last value for adaptive solvers) */
    double   eig;        /* coarse eigenvalue estimate that produced it (0 if rho was given) */
    double   lambda0;    /* max |X'y| on the standardised scale */
    b200admm_timing t;
} b200admm_path;

/* admm_lasso(x, y, lambda, nlambda, lmin_ratio, standardize, intercept, opts)
 * lambda_given: user lambdas sorted decreasingly by the R side (may be NULL),
 * nlambda_given < 1 -> log-linear grid of `nlambda` values down to lmin_ratio * lambda_max; lmin_ratio <= 0 selects the
 * front end's default (R/30_admm_lasso.R:44: 0.01 if n < p else 1e-4) from the GLOBAL row count of a row-sharded run. */
int b200admm_lasso(const b200admm_data* d, const double* lambda_given, int nlambda_given,
                   int nlambda, double lmin_ratio, int standardize, int intercept,
                   const b200admm_opts* opts, b200admm_path* out);

/* admm_enet(..., alpha, opts) */
int b200admm_enet(const b200admm_data* d, const double* lambda_given, int nlambda_given,
                  int nlambda, double lmin_ratio, int standardize, int intercept,
                  double alpha, const b200admm_opts* opts, b200admm_path* out);

/* admm_parlasso(..., nthread, opts): row-split consensus over `nthread` blocks.  In this
 * library a block is a GPU: nthread must equal the size of the communicator installed with
 * b200admm_comm_init (one process per GPU), or nthread blocks are run on the current device
 * when no communicator is installed.  `d` holds the caller's rows only when a communicator
 * is installed (see b200admm_comm_init), the whole matrix otherwise. */
int b200admm_parlasso(const b200admm_data* d, const double* lambda_given, int nlambda_given,
                      int nlambda, double lmin_ratio, int standardize, int intercept,
                      int nthread, const b200admm_opts* opts, b200admm_path* out);

void b200admm_free_path(b200admm_path* out);

/* admm_lad(x, y, intercept, opts) -> List(beta = double[p + 1], niter) (src/LAD.cpp:44-45) */
typedef struct b200admm_dense {
    int64_t len;
    double* beta;
    int     niter;
    double  rho;
    b200admm_timing t;
} b200admm_dense;
int  b200admm_lad(const b200admm_data* d, int intercept, const b200admm_opts* opts, b200admm_dense* out);
void b200admm_free_dense(b200admm_dense* out);

/* admm_bp(x, y, opts) -> List(beta = dgCMatrix p x 1, niter) (src/BP.cpp:38-43).
 * Returned as a one-column b200admm_path with nrow = p (no intercept row). */
int b200admm_bp(const b200admm_data* d, const b200admm_opts* opts, b200admm_path* out);

/* ---- row-sharded runs (one process per GPU) ------------------------------------------------
 * b200admm_comm_id writes a B200ADMM_COMM_ID_BYTES identifier on one rank; the host program
 * broadcasts it and every rank calls b200admm_comm_init.  After that b200admm_lasso/_enet
 * treat `d` as this rank's row block of a taller matrix (global standardisation and Gram via
 * all-reduce; the factorisation is replicated and the iterations are sharded over the rows of K^-1
 * with the exchange fused into the persistent kernel over NVLink peer memory -- every rank returns
 * the identical result, equal to the single-GPU solver's up to summation order), and
 * b200admm_parlasso runs the reference's consensus algorithm with one block per rank
 * (src/PADMMLasso.h:163-178) and one all-reduce per iteration. */
#define B200ADMM_COMM_ID_BYTES 128
int  b200admm_comm_id(void* id_out);
int  b200admm_comm_init(const void* id, int rank, int nranks);
void b200admm_comm_destroy(void);
/* on != 0: until switched off again every call treats `d` as a whole matrix on this GPU alone, exactly as if no
 * communicator were installed (bench.py / the multi-GPU tests fit the undivided problem on one rank to check
 * the sharded result against it).  Not collective. */
void b200admm_comm_suspend(int on);

/* ---- utilities ---------------------------------------------------------------------------- */
const char* b200admm_last_error(void);
int  b200admm_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long b200admm_launch_count(void);
/* device seconds the Gram-matrix kernel(s) of the most recent lasso / enet call took (CUDA events around
 * the tensor-core launches only; the `gram` field of b200admm_timing also covers X'y, the operand split
 * and, for host input, the overlapped copy).  bench.py's tensor roofline is computed from it. */
double b200admm_last_gram_seconds(void);
/* Work counters of the most recent wide (n <= p) lasso / enet call, whose per-iteration traffic depends on the
 * support: out4 = {algorithmic bytes of the iteration phase (sum over iterations of 4 n p for a regular step or
 * 4 n nnz_k for an active-set step, + 4 n nnz_{k+1} for the z step + 48 n of vector traffic), number of regular
 * steps, number of active-set steps, 0}.  bench.py's HBM roofline of that path is computed from it. */
void b200admm_last_work(double* out4);
/* the cudaStream_t every kernel of this library is launched on (for event timing by the host
 * program); NULL if no device is usable */
void* b200admm_stream(void);
/* Device blocks of 64 MB and more (the float32 working copy of x, the p x p matrices) are kept for
 * reuse by the next call instead of being returned to the driver; this frees them.
 * Environment B200ADMM_CACHE=0 disables the recycling altogether. */
void b200admm_release_cache(void);
/* name, SM count and total memory of the current device; returns B200ADMM_ENODEVICE if none */
int  b200admm_device_info(char* name, int name_len, int* sm_count, int64_t* mem_bytes);

/* Per-iteration trace of one lambda (parity tests): rows of 5 doubles
 * {eps_primal, resid_primal, eps_dual, resid_dual, rho}.  Set before a call; the library
 * fills up to `cap` rows for lambda index `which` and stores the row count in *nrows. */
void b200admm_set_trace(double* buf, int cap, int which, int* nrows);
/* Parity checks at sizes where X cannot be handed to a CPU program: the next b200admm_lasso / _enet call with
 * n > p copies the Gram matrix X'X of the standardised data (float, p x p column-major, full symmetric, BEFORE
 * rho is added) to gram_host, X'y (p floats) to xy_host and DataStd's statistics to stats_host (2 p + 2 floats:
 * meanX[p], scaleX[p], meanY, scaleY); any may be NULL, all NULL switches it off.
 * The copies add device->host traffic to that call: do not time it. */
void b200admm_set_capture(float* gram_host, float* xy_host, float* stats_host);

/* Synthetic design of the reference's benchmarks (README.md:195-201): X_ij ~ N(mean_x, sd_x^2)
 * i.i.d. (counter-based Philox4x32-10, so any row block can be produced independently),
 * beta* = nsig leading U(0,1) entries, y = X beta* + N(0, noise^2).  Writes float32 into device
 * memory: x_dev (nrows x p, column-major, ld = nrows) holds global rows [row0, row0 + nrows). */
int b200admm_synth_f32(void* x_dev, void* y_dev, int64_t nrows, int64_t p, int64_t row0,
                       uint64_t seed, float mean_x, float sd_x, int nsig, float noise);

/* ---- kernel-level entry points (microbenchmarks, unit tests, ncu targets) ------------------
 * All pointers are device pointers on the current device unless named *_host. */
int b200admm_k_standardize_f32(const void* x_in, void* x_out, void* y_inout, int64_t n, int64_t p,
                               int standardize, int intercept,
                               float* meanx_host, float* scalex_host, float* meany_scaley_host);
/* use_tensor: 0 CUDA cores, 1 tcgen05 3xTF32 with truncation split (raw tile serves as hi), 2 the same with a
 * round-to-nearest split (any fp32 data), 3 tcgen05 3xFP16 hi/lo split (unit-scale columns only, i.e. data
 * standardised by DataStd -- what the solvers use there; values beyond the fp16 range are an error here) */
int b200admm_k_gram_f32(const void* x, int64_t n, int64_t p, void* g /* p x p, full */, int use_tensor);
/* Host-only planners (no device needed; unit-tested on CPU): the work distribution of the fp16 Gram kernel --
 * cover[tile * nk + stage] = number of CTA pairs computing that 32-row stage of that 256 x 256 tile (must be 1
 * everywhere), per_pair[q] = stages handled by pair q, *nslices = canonical K-slices per tile, *split_tiles =
 * tiles of the last round that are cut along K -- and the column panels of the pipelined host ingest
 * (returns the number of panels, begin[0 .. npanels] their first columns; -1 on bad arguments). */
int b200admm_k_gram_plan(int ntiles, int npairs, int nk, int* cover, long long* per_pair, int* nslices, int* split_tiles);
int b200admm_k_panel_schedule(int64_t p, int64_t panel_cols, int64_t* begin, int cap);
/* Host-only replay of the row assignment of the single-GPU iteration kernel that reads one triangle of K^-1
 * (fadmm_tall.cu: tall_path_tri_kernel): CTA c owns the rows [rows[4c], rows[4c+1]) and [rows[4c+2], rows[4c+3]) of a p x p
 * matrix on a device with `sms` SMs.  Returns the grid size, 0 when the shape does not take that kernel (shared memory),
 * -1 on bad arguments / cap too small; smem_bytes (optional) = dynamic shared memory of the launch. */
int b200admm_k_tri_plan(int p, int sms, int* rows, int cap, long long* smem_bytes);
/* Host-only: the default lambda sequence of admm_lasso / admm_enet (src/Lasso.cpp:78-89,
 * `lambda.setLinSpaced(nlambda, log(lmax), log(lmin)).exp()` with Eigen's LinSpaced semantics: a single
 * value is the HIGH end, i.e. nlambda = 1 fits at lmin_ratio * lmax).  out: nlambda doubles. */
int b200admm_k_lambda_grid(double lmax, double lmin_ratio, int nlambda, double* out);
int b200admm_k_gemv_t_f32(const void* a, int64_t m, int64_t ncol, const void* v, void* out);
/* The coarse lambda_max estimate behind the default rho / gamma (src/ADMMLassoTall.h:194-202: Spectra
 * SymEigsSolver<float, LARGEST_ALGE>(op, 1, 3).compute(10, 0.1)) of the symmetric float32 matrix s (n x n, full
 * storage, device).  info_host (optional, 3 ints): products with s, restarts, converged flag. */
int b200admm_k_coarse_eig_f32(const void* s, int64_t n, float* ev_host, int* info_host);
/* C (M x N, ldc) op= A' B on the tcgen05 3xTF32 CTA-pair kernel; A: K x M, B: K x N, column-major, device.
 * tile_mode 0 all tiles / 1 on and below / 2 on and above the diagonal (256 x 256 tiles); klo_mode, khi_mode: K range
 * per tile in 256-row blocks (0 all; 1 block I; 2 block J; 3 max(I, J) resp. min(I, J)); epi 0 C = A'B, 1 C -= A'B,
 * 2 C = -A'B.  The building block of the blocked factorisation (kernels.h: gemm_tn_tensor). */
int b200admm_k_gemm_tn_f32(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                           void* c, int64_t ldc, int tile_mode, int klo_mode, int khi_mode, int epi);
/* C (m x n, ldc) = alpha op(A) op(B) + beta C in float64 (device, column-major): the product behind the LAD / BP Gram
 * matrices and triangular solves (dsyrk_ / dtrsm_ in src/ADMMLAD.h:186-201, src/ADMMBP.h:167-182).  mode: bit 0 store i >= j
 * only, bit 1 mirror, bits 2-5 triangular operands (kernels.h: GEMM_*).  Tensor-core (mma.sync f64) tiles when the
 * 128 x 128 grid fills the chip, CUDA-core tiles otherwise.  ms_out (optional): kernel time of `repeats` launches / repeats. */
int b200admm_k_gemm_f64(int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha, const void* a, int64_t lda,
                        const void* b, int64_t ldb, double beta, void* c, int64_t ldc, int mode, float* ms_out, int repeats);
int b200admm_k_chol_f32(void* a, int64_t p, int* info_host);                 /* lower, in place */
int b200admm_k_spd_inverse_f32(void* a, int64_t p, void* work, int* info_host); /* a <- a^-1 (full) */
/* fused z + u + residual + norms pass of the accelerated loop on vectors of length len
 * (src/ADMMLassoTall.h:81-95, src/FADMMBase.h:195-211): reads x, adj_y, old_z, adj_z; writes z, y;
 * sums_host[6] = {|r|^2, |z-old_z|^2, |z-adj_z|^2, |x|^2, |z|^2, |y|^2}.  Returns kernel ms. */
int b200admm_k_fused_zu_f32(const void* x, const void* adj_y, const void* old_z, const void* adj_z,
                            void* z, void* y, int64_t len, double lambda, double rho,
                            int enet, double alpha, double* sums_host, float* ms_out, int repeats);

#ifdef __cplusplus
}
#endif
#endif /* B200ADMM_H */
