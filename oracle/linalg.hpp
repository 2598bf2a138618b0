// oracle/linalg.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Dense column-major helpers for the CPU oracle.  The reference gets these from
// Eigen (float paths) and R's BLAS (double paths); neither is available in this
// image, so they are restated here as plain loops (OpenMP where the work is big)
// with an optional fast path through OpenBLAS reached by dlopen (the scipy wheel's
// libscipy_openblas, symbols prefixed `scipy_`).  The BLAS path is what the CPU
// baseline in bench.py times -- it corresponds to the reference's *recommended*
// build (drop NO_FLOAT_BLAS, link OpenBLAS; /root/reference/README.md:26-38).
//
// Reference call sites restated:
//   cross_prod_lower   /root/reference/src/Linalg/BlasWrapper.h:73-112   -> gram_tn_lower
//   tcross_prod_lower  /root/reference/src/Linalg/BlasWrapper.h:115-154  -> gram_nt_lower
//   mat_vec_prod/tprod /root/reference/src/Linalg/BlasWrapper.h:46-66    -> gemv_n / gemv_t
//   Eigen::LLT compute/solve (ADMMLassoTall.h:79,205 etc.)                -> chol_lower / chol_solve
//   dtrsm_ (ADMMLAD.h:198-199, ADMMBP.h:181-182)                          -> trsm_*
//   dsymv_ (ADMMLAD.h:72-73)                                              -> symv_lower
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include <algorithm>
#include <dlfcn.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

typedef long long i64;

// ---------------------------------------------------------------- BLAS hooks
struct BlasTable {
    void* handle = nullptr;
    // Fortran ABI, LP64
    void (*ssyrk)(const char*, const char*, const int*, const int*, const float*, const float*, const int*, const float*, float*, const int*) = nullptr;
    void (*dsyrk)(const char*, const char*, const int*, const int*, const double*, const double*, const int*, const double*, double*, const int*) = nullptr;
    void (*spotrf)(const char*, const int*, float*, const int*, int*) = nullptr;
    void (*dpotrf)(const char*, const int*, double*, const int*, int*) = nullptr;
    void (*strsv)(const char*, const char*, const char*, const int*, const float*, const int*, float*, const int*) = nullptr;
    void (*dtrsv)(const char*, const char*, const char*, const int*, const double*, const int*, double*, const int*) = nullptr;
    void (*sgemv)(const char*, const int*, const int*, const float*, const float*, const int*, const float*, const int*, const float*, float*, const int*) = nullptr;
    void (*dgemv)(const char*, const int*, const int*, const double*, const double*, const int*, const double*, const int*, const double*, double*, const int*) = nullptr;
    void (*ssymv)(const char*, const int*, const float*, const float*, const int*, const float*, const int*, const float*, float*, const int*) = nullptr;
    void (*dsymv)(const char*, const int*, const double*, const double*, const int*, const double*, const int*, const double*, double*, const int*) = nullptr;
    void (*dtrsm)(const char*, const char*, const char*, const char*, const int*, const int*, const double*, const double*, const int*, double*, const int*) = nullptr;
    void (*set_threads)(int) = nullptr;
    int  (*get_threads)() = nullptr;
    bool on = false;
};

inline BlasTable& blas() { static BlasTable t; return t; }

inline int blas_load(const char* path)
{
    BlasTable& t = blas();
    if (t.handle) return 0;
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    auto sym = [&](const char* base) -> void* {
        // try scipy_ prefix first, then plain
        std::string a = std::string("scipy_") + base;
        void* p = dlsym(h, a.c_str());
        if (!p) p = dlsym(h, base);
        return p;
    };
    t.ssyrk  = (decltype(t.ssyrk))  sym("ssyrk_");
    t.dsyrk  = (decltype(t.dsyrk))  sym("dsyrk_");
    t.spotrf = (decltype(t.spotrf)) sym("spotrf_");
    t.dpotrf = (decltype(t.dpotrf)) sym("dpotrf_");
    t.strsv  = (decltype(t.strsv))  sym("strsv_");
    t.dtrsv  = (decltype(t.dtrsv))  sym("dtrsv_");
    t.sgemv  = (decltype(t.sgemv))  sym("sgemv_");
    t.dgemv  = (decltype(t.dgemv))  sym("dgemv_");
    t.ssymv  = (decltype(t.ssymv))  sym("ssymv_");
    t.dsymv  = (decltype(t.dsymv))  sym("dsymv_");
    t.dtrsm  = (decltype(t.dtrsm))  sym("dtrsm_");
    t.set_threads = (decltype(t.set_threads)) sym("openblas_set_num_threads");
    t.get_threads = (decltype(t.get_threads)) sym("openblas_get_num_threads");
    if (!t.ssyrk || !t.dsyrk || !t.spotrf || !t.dpotrf || !t.strsv || !t.dtrsv ||
        !t.sgemv || !t.dgemv || !t.ssymv || !t.dsymv || !t.dtrsm) {
        dlclose(h);
        t = BlasTable();
        return -2;
    }
    t.handle = h;
    t.on = true;
    return 0;
}

inline bool fits_int(i64 a) { return a < 2147483647LL; }

// ------------------------------------------------------------ level-1 pieces
template <class T> inline T dot(const T* a, const T* b, i64 n)
{
    // accumulate in the vector's own scalar type, like Eigen's float/double reductions
    T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    i64 i = 0;
    for (; i + 4 <= n; i += 4) {
        s0 += a[i] * b[i]; s1 += a[i + 1] * b[i + 1];
        s2 += a[i + 2] * b[i + 2]; s3 += a[i + 3] * b[i + 3];
    }
    for (; i < n; i++) s0 += a[i] * b[i];
    return (s0 + s1) + (s2 + s3);
}
template <class T> inline T sqnorm(const T* a, i64 n) { return dot(a, a, n); }
template <class T> inline T norm2(const T* a, i64 n) { return std::sqrt(sqnorm(a, n)); }
template <class T> inline T mean(const T* a, i64 n)
{
    T s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    i64 i = 0;
    for (; i + 4 <= n; i += 4) { s0 += a[i]; s1 += a[i + 1]; s2 += a[i + 2]; s3 += a[i + 3]; }
    for (; i < n; i++) s0 += a[i];
    return ((s0 + s1) + (s2 + s3)) / T(n);
}
// strictly sequential sum of squares -- Eigen's SparseVector reductions and the
// reference's hand-written diff_squared_norm loops accumulate one entry at a time
template <class T> inline T sqnorm_seq(const T* a, i64 n)
{
    T r = 0;
    for (i64 i = 0; i < n; i++) r += a[i] * a[i];
    return r;
}
template <class T> inline T diff_sqnorm_seq(const T* a, const T* b, i64 n)
{
    T r = 0;
    for (i64 i = 0; i < n; i++) { T v = a[i] - b[i]; r += v * v; }
    return r;
}

// ------------------------------------------------------------------- level-2
// y = A x, A is m x n column-major (lda = m)
template <class T> void gemv_n(const T* A, i64 m, i64 n, const T* x, T* y);
// y = A^T x
template <class T> void gemv_t(const T* A, i64 m, i64 n, const T* x, T* y);

template <class T> inline void gemv_n_plain(const T* A, i64 m, i64 n, const T* x, T* y)
{
    std::fill(y, y + m, T(0));
#pragma omp parallel if (m * n > 200000)
    {
        // split rows between threads so every thread owns a slice of y
        int nt = 1, id = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads(); id = omp_get_thread_num();
#endif
        i64 r0 = m * id / nt, r1 = m * (id + 1) / nt;
        for (i64 j = 0; j < n; j++) {
            const T xj = x[j];
            const T* a = A + j * m;
            for (i64 i = r0; i < r1; i++) y[i] += a[i] * xj;
        }
    }
}
template <class T> inline void gemv_t_plain(const T* A, i64 m, i64 n, const T* x, T* y)
{
#pragma omp parallel for schedule(static) if (m * n > 200000)
    for (i64 j = 0; j < n; j++) y[j] = dot(A + j * m, x, m);
}
template <> inline void gemv_n<float>(const float* A, i64 m, i64 n, const float* x, float* y)
{
    if (blas().on && fits_int(m) && fits_int(n)) {
        int M = (int)m, N = (int)n, one = 1; float al = 1.f, be = 0.f;
        blas().sgemv("N", &M, &N, &al, A, &M, x, &one, &be, y, &one);
    } else gemv_n_plain(A, m, n, x, y);
}
template <> inline void gemv_n<double>(const double* A, i64 m, i64 n, const double* x, double* y)
{
    if (blas().on && fits_int(m) && fits_int(n)) {
        int M = (int)m, N = (int)n, one = 1; double al = 1., be = 0.;
        blas().dgemv("N", &M, &N, &al, A, &M, x, &one, &be, y, &one);
    } else gemv_n_plain(A, m, n, x, y);
}
template <> inline void gemv_t<float>(const float* A, i64 m, i64 n, const float* x, float* y)
{
    if (blas().on && fits_int(m) && fits_int(n)) {
        int M = (int)m, N = (int)n, one = 1; float al = 1.f, be = 0.f;
        blas().sgemv("T", &M, &N, &al, A, &M, x, &one, &be, y, &one);
    } else gemv_t_plain(A, m, n, x, y);
}
template <> inline void gemv_t<double>(const double* A, i64 m, i64 n, const double* x, double* y)
{
    if (blas().on && fits_int(m) && fits_int(n)) {
        int M = (int)m, N = (int)n, one = 1; double al = 1., be = 0.;
        blas().dgemv("T", &M, &N, &al, A, &M, x, &one, &be, y, &one);
    } else gemv_t_plain(A, m, n, x, y);
}

// y = S x with S symmetric, only the lower triangle of the n x n array is read
template <class T> inline void symv_lower_plain(const T* S, i64 n, const T* x, T* y)
{
    std::fill(y, y + n, T(0));
    for (i64 j = 0; j < n; j++) {
        const T* c = S + j * n;
        const T xj = x[j];
        T acc = c[j] * xj;
        for (i64 i = j + 1; i < n; i++) { y[i] += c[i] * xj; acc += c[i] * x[i]; }
        y[j] += acc;
    }
}
template <class T> void symv_lower(const T* S, i64 n, const T* x, T* y);
template <> inline void symv_lower<float>(const float* S, i64 n, const float* x, float* y)
{
    if (blas().on && fits_int(n)) {
        int N = (int)n, one = 1; float al = 1.f, be = 0.f;
        blas().ssymv("L", &N, &al, S, &N, x, &one, &be, y, &one);
    } else symv_lower_plain(S, n, x, y);
}
template <> inline void symv_lower<double>(const double* S, i64 n, const double* x, double* y)
{
    if (blas().on && fits_int(n)) {
        int N = (int)n, one = 1; double al = 1., be = 0.;
        blas().dsymv("L", &N, &al, S, &N, x, &one, &be, y, &one);
    } else symv_lower_plain(S, n, x, y);
}

// ------------------------------------------------------------------- level-3
// G(p x p, lower) = X^T X, X is n x p column-major
template <class T> inline void gram_tn_lower_plain(const T* X, i64 n, i64 p, T* G)
{
    // row-chunked so that a chunk of all p columns stays in cache
    std::fill(G, G + p * p, T(0));
    const i64 RB = 2048;
    for (i64 r0 = 0; r0 < n; r0 += RB) {
        const i64 rb = std::min(RB, n - r0);
#pragma omp parallel for schedule(dynamic, 4) if (p * rb > 50000)
        for (i64 j = 0; j < p; j++) {
            const T* xj = X + j * n + r0;
            for (i64 i = j; i < p; i++) G[j * p + i] += dot(X + i * n + r0, xj, rb);
        }
    }
}
// G(n x n, lower) = X X^T
template <class T> inline void gram_nt_lower_plain(const T* X, i64 n, i64 p, T* G)
{
    std::fill(G, G + n * n, T(0));
#pragma omp parallel for schedule(dynamic, 4) if (n * p > 50000)
    for (i64 j = 0; j < n; j++) {
        for (i64 k = 0; k < p; k++) {
            const T* c = X + k * n;
            const T v = c[j];
            T* g = G + j * n;
            for (i64 i = j; i < n; i++) g[i] += c[i] * v;
        }
    }
}
template <class T> void gram_tn_lower(const T* X, i64 n, i64 p, T* G);
template <class T> void gram_nt_lower(const T* X, i64 n, i64 p, T* G);
template <> inline void gram_tn_lower<float>(const float* X, i64 n, i64 p, float* G)
{
    if (blas().on && fits_int(n) && fits_int(p)) {
        int N = (int)p, K = (int)n; float al = 1.f, be = 0.f;
        std::fill(G, G + p * p, 0.f);
        blas().ssyrk("L", "T", &N, &K, &al, X, &K, &be, G, &N);
    } else gram_tn_lower_plain(X, n, p, G);
}
template <> inline void gram_tn_lower<double>(const double* X, i64 n, i64 p, double* G)
{
    if (blas().on && fits_int(n) && fits_int(p)) {
        int N = (int)p, K = (int)n; double al = 1., be = 0.;
        std::fill(G, G + p * p, 0.);
        blas().dsyrk("L", "T", &N, &K, &al, X, &K, &be, G, &N);
    } else gram_tn_lower_plain(X, n, p, G);
}
template <> inline void gram_nt_lower<float>(const float* X, i64 n, i64 p, float* G)
{
    if (blas().on && fits_int(n) && fits_int(p)) {
        int N = (int)n, K = (int)p; float al = 1.f, be = 0.f;
        std::fill(G, G + n * n, 0.f);
        blas().ssyrk("L", "N", &N, &K, &al, X, &N, &be, G, &N);
    } else gram_nt_lower_plain(X, n, p, G);
}
template <> inline void gram_nt_lower<double>(const double* X, i64 n, i64 p, double* G)
{
    if (blas().on && fits_int(n) && fits_int(p)) {
        int N = (int)n, K = (int)p; double al = 1., be = 0.;
        std::fill(G, G + n * n, 0.);
        blas().dsyrk("L", "N", &N, &K, &al, X, &N, &be, G, &N);
    } else gram_nt_lower_plain(X, n, p, G);
}

// In-place lower Cholesky of the n x n column-major array A (upper part untouched).
// Returns 0 on success, j+1 if the leading minor of order j+1 is not positive.
template <class T> inline int chol_lower_plain(T* A, i64 n)
{
    // left-looking, column at a time; columns are contiguous
    std::vector<T> col(n);
    for (i64 j = 0; j < n; j++) {
        T* aj = A + j * n;
        // aj[j:] -= sum_k L[j,k] * L[j:,k]
        for (i64 k = 0; k < j; k++) {
            const T* lk = A + k * n;
            const T ljk = lk[j];
            if (ljk == T(0)) continue;
            for (i64 i = j; i < n; i++) aj[i] -= lk[i] * ljk;
        }
        if (!(aj[j] > T(0))) return (int)(j + 1);
        const T d = std::sqrt(aj[j]);
        aj[j] = d;
        for (i64 i = j + 1; i < n; i++) aj[i] /= d;
    }
    return 0;
}
template <class T> int chol_lower(T* A, i64 n);
template <> inline int chol_lower<float>(float* A, i64 n)
{
    if (blas().on && fits_int(n)) { int N = (int)n, info = 0; blas().spotrf("L", &N, A, &N, &info); return info; }
    return chol_lower_plain(A, n);
}
template <> inline int chol_lower<double>(double* A, i64 n)
{
    if (blas().on && fits_int(n)) { int N = (int)n, info = 0; blas().dpotrf("L", &N, A, &N, &info); return info; }
    return chol_lower_plain(A, n);
}

// b <- L^{-1} b  (forward), column-oriented
template <class T> inline void trsv_lower_n_plain(const T* L, i64 n, T* b)
{
    for (i64 j = 0; j < n; j++) {
        const T* c = L + j * n;
        const T v = b[j] / c[j];
        b[j] = v;
        for (i64 i = j + 1; i < n; i++) b[i] -= c[i] * v;
    }
}
// b <- L^{-T} b (backward), uses columns of L as rows of L^T
template <class T> inline void trsv_lower_t_plain(const T* L, i64 n, T* b)
{
    for (i64 j = n - 1; j >= 0; j--) {
        const T* c = L + j * n;
        T s = b[j];
        for (i64 i = j + 1; i < n; i++) s -= c[i] * b[i];
        b[j] = s / c[j];
    }
}
template <class T> void trsv_lower_n(const T* L, i64 n, T* b);
template <class T> void trsv_lower_t(const T* L, i64 n, T* b);
template <> inline void trsv_lower_n<float>(const float* L, i64 n, float* b)
{
    if (blas().on && fits_int(n)) { int N = (int)n, one = 1; blas().strsv("L", "N", "N", &N, L, &N, b, &one); }
    else trsv_lower_n_plain(L, n, b);
}
template <> inline void trsv_lower_t<float>(const float* L, i64 n, float* b)
{
    if (blas().on && fits_int(n)) { int N = (int)n, one = 1; blas().strsv("L", "T", "N", &N, L, &N, b, &one); }
    else trsv_lower_t_plain(L, n, b);
}
template <> inline void trsv_lower_n<double>(const double* L, i64 n, double* b)
{
    if (blas().on && fits_int(n)) { int N = (int)n, one = 1; blas().dtrsv("L", "N", "N", &N, L, &N, b, &one); }
    else trsv_lower_n_plain(L, n, b);
}
template <> inline void trsv_lower_t<double>(const double* L, i64 n, double* b)
{
    if (blas().on && fits_int(n)) { int N = (int)n, one = 1; blas().dtrsv("L", "T", "N", &N, L, &N, b, &one); }
    else trsv_lower_t_plain(L, n, b);
}
// b <- (L L^T)^{-1} b
template <class T> inline void chol_solve(const T* L, i64 n, T* b)
{
    trsv_lower_n(L, n, b);
    trsv_lower_t(L, n, b);
}

// B (m x k) <- B L^{-T}   ("R","L","T","N"), L is k x k lower
inline void trsm_right_lower_trans(const double* L, i64 k, double* B, i64 m)
{
    if (blas().on && fits_int(k) && fits_int(m)) {
        int M = (int)m, K = (int)k; double al = 1.;
        blas().dtrsm("R", "L", "T", "N", &M, &K, &al, L, &K, B, &M);
        return;
    }
    // column j of result: (B_j - sum_{c<j} R_c L[j,c]) / L[j,j]
    for (i64 j = 0; j < k; j++) {
        double* bj = B + j * m;
        for (i64 c = 0; c < j; c++) {
            const double l = L[c * k + j];
            const double* rc = B + c * m;
            for (i64 i = 0; i < m; i++) bj[i] -= rc[i] * l;
        }
        const double d = L[j * k + j];
        for (i64 i = 0; i < m; i++) bj[i] /= d;
    }
}
// B (k x m) <- L^{-1} B   ("L","L","N","N"), L is k x k lower
inline void trsm_left_lower_notrans(const double* L, i64 k, double* B, i64 m)
{
    if (blas().on && fits_int(k) && fits_int(m)) {
        int M = (int)k, N = (int)m; double al = 1.;
        blas().dtrsm("L", "L", "N", "N", &M, &N, &al, L, &M, B, &M);
        return;
    }
#pragma omp parallel for schedule(static) if (k * m > 50000)
    for (i64 c = 0; c < m; c++) trsv_lower_n_plain(L, k, B + c * k);
}

}  // namespace oracle
