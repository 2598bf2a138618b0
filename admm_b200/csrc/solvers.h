// solvers.h -- host-side path drivers shared between the C ABI (capi.cu) and the solver files.
#pragma once
#include "../../include/b200admm.h"
#include "common.cuh"
#include <vector>

namespace b200 {

struct Context {
    bool ready = false;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy = nullptr;     // second stream for host->device copies overlapped with compute
};
Context& ctx();
cudaStream_t copy_stream();                      // throws CodeError(B200ADMM_ENODEVICE) when no B200 is usable

struct TraceRequest {
    double* buf = nullptr;
    int cap = 0;
    int which = -1;
    int* nrows = nullptr;
};
TraceRequest& trace_request();

// b200admm_set_capture: host buffers that receive the Gram matrix X'X of the standardised data (before rho is
// added; p x p, column-major, full symmetric) and X'y of the next tall lasso / enet call (parity checks)
struct CaptureRequest {
    float* gram = nullptr;
    float* xy = nullptr;
    float* stats = nullptr;       // 2 p + 2 floats: meanX[p], scaleX[p], meanY, scaleY (DataStd)
};
CaptureRequest& capture_request();

double wall_now();

struct LassoRequest {
    const b200admm_data* d;
    const double* lambda_given;
    int nlambda_given;
    int nlambda;
    double lmin_ratio;
    bool standardize, intercept;
    bool enet;
    double alpha;
    b200admm_opts opts;
};

// capi.cu helpers
void make_lambda_grid(double lmax, double ratio, int nl, std::vector<double>& out);
void free_path(b200admm_path* out);
template <class T>
void assemble_csc(const std::vector<std::vector<T>>& cols, const std::vector<T>& beta0, bool with_intercept,
                  i64 p, b200admm_path* out);
template <class T>
T recover_sparse(int flag, std::vector<T>& coef, const std::vector<T>& meanX, const std::vector<T>& scaleX, T meanY, T scaleY);
std::vector<i64> host_panel_schedule(i64 p, i64 pw);
void ingest_f32(cudaStream_t s, const void* src, int dtype, size_t count, float* dst);
void ingest_f64(cudaStream_t s, const void* src, int dtype, size_t count, double* dst);
void add_to_diagonal(cudaStream_t s, float* A, i64 ld, i64 p, float v);
float coarse_eig_device(cudaStream_t s, const float* S, i64 n, i64 lds, int* nmatvec);
void spd_inverse_f32(cudaStream_t s, float* a, i64 p, i64 ld, float* W, int* info_host);
template <class T> void spd_inverse(cudaStream_t s, T* a, i64 p, i64 ld, T* W, int* info_host, T* keep_factor);

// stdize.cu pieces used by the drivers
template <class T> void mean_from_sums(cudaStream_t s, const T* sums, i64 p, i64 n_total, T* mean);
template <class T> void scale_from_sumsq(cudaStream_t s, const T* sumsq, i64 p, i64 n_total, bool sd_form, T* scale, T* inv);
template <class T> void column_apply(cudaStream_t s, const T* Xin, T* Xout, i64 n, i64 p, i64 ld, i64 ld_out, const T* mean, const T* factor, const T* divisor);

// solver_tall.cu / solver_wide.cu : admm_lasso / admm_enet (dispatch on n > p, Lasso.cpp:73-76)
void solve_lasso_like(const LassoRequest& rq, b200admm_path* out);
void solve_wide(const LassoRequest& rq, b200admm_path* out);
// solver_consensus.cu : admm_parlasso
void solve_consensus(const LassoRequest& rq, int nthread, b200admm_path* out);
// solver_lad_bp.cu
void solve_lad(const b200admm_data* d, bool intercept, const b200admm_opts& o, b200admm_dense* out);
void solve_bp(const b200admm_data* d, const b200admm_opts& o, b200admm_path* out);

}  // namespace b200
