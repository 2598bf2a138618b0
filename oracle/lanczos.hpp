// oracle/lanczos.hpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Restatement of the *coarse* largest-eigenvalue estimate the reference obtains from
// its vendored Spectra:  SymEigsSolver<Scalar, LARGEST_ALGE>(op, nev=1, ncv=3),
// init() with the deterministic LCG start vector, compute(maxit=10, tol=0.1).
//
// Pinned independently of the product: the library's own Lanczos driver is a sibling of this file (same reading
// of Spectra, written twice), so this restatement is checked against a third, structurally different one --
// tests/spectra_numpy.py (matrix form, LAPACK for the 3 x 3 eigenproblem) -- in
// tests/test_lanczos_independent_cpu.py: 25 matrices, 0-3 implicit restarts, 5e-6 relative.
//
// Follows (by reading, not by copying):
//   /root/reference/src/Spectra/SymEigsSolver.h:201-280  (Lanczos step + re-orthogonalisation)
//   /root/reference/src/Spectra/SymEigsSolver.h:283-323  (implicit restart)
//   /root/reference/src/Spectra/SymEigsSolver.h:326-397  (convergence test, adjusted nev, Ritz pairs)
//   /root/reference/src/Spectra/SymEigsSolver.h:494-587  (init, compute)
//   /root/reference/src/Spectra/SimpleRandom.h:38-76     (Lehmer generator, Schrage form)
//   /root/reference/src/Spectra/LinAlg/UpperHessenbergQR.h:415-602 (tridiagonal QR, RQ, Y*Q)
//   /root/reference/src/Spectra/LinAlg/TridiagEigen.h:43-170       (implicit symmetric QR)
// Call sites: /root/reference/src/ADMMLassoTall.h:194-201, ADMMLassoWide.h:200-207.
//
// Everything is done in Scalar (float for the lasso/enet paths), as in the reference,
// because the iterate path of the solver depends on the *rounded, unconverged* value.
#pragma once
#include "linalg.hpp"
#include <limits>

namespace oracle {

// Park-Miller minimal standard generator; state advanced before each draw; output in [-0.5, 0.5)
struct LehmerStream {
    long state;
    explicit LehmerStream(unsigned long seed) : state(seed ? (long)(seed & 2147483647UL) : 1L) {}
    long next()
    {
        // 16807 * state mod (2^31 - 1), computed without overflow of 32-bit halves
        const long M = 2147483647L;
        unsigned long lo = 16807UL * (unsigned long)(state & 0xFFFF);
        unsigned long hi = 16807UL * ((unsigned long)state >> 16);
        lo += (hi & 0x7FFF) << 16;
        if ((long)lo > M) { lo &= (unsigned long)M; ++lo; }
        lo += hi >> 15;
        if ((long)lo > M) { lo &= (unsigned long)M; ++lo; }
        state = (long)lo;
        return state;
    }
    template <class T> void fill(T* v, i64 n)
    {
        for (i64 i = 0; i < n; i++) v[i] = T(next()) / T(2147483647L) - T(0.5);
    }
};

template <class T> struct GivensPair { T c, s; };

// Eigen-style Givens: [c s; -s c]^T [p; q] = [r; 0]
template <class T> inline GivensPair<T> make_givens(T p, T q)
{
    GivensPair<T> g;
    if (q == T(0)) { g.c = p < T(0) ? T(-1) : T(1); g.s = T(0); }
    else if (p == T(0)) { g.c = T(0); g.s = q < T(0) ? T(1) : T(-1); }
    else if (std::abs(p) > std::abs(q)) {
        T t = q / p, u = std::sqrt(T(1) + t * t);
        if (p < T(0)) u = -u;
        g.c = T(1) / u; g.s = -t * g.c;
    } else {
        T t = p / q, u = std::sqrt(T(1) + t * t);
        if (q < T(0)) u = -u;
        g.s = -T(1) / u; g.c = -t * g.s;
    }
    return g;
}

// Eigen decomposition of a small symmetric tridiagonal matrix (m <= 8) by implicit QR
// with Wilkinson shift.  d: diagonal (m) -> eigenvalues; e: sub-diagonal (m-1);
// Q (m x m, column-major) -> eigenvectors.  Returns false if it does not finish.
template <class T> inline bool tridiag_eig_small(int m, T* d, T* e, T* Q)
{
    for (int i = 0; i < m * m; i++) Q[i] = T(0);
    for (int i = 0; i < m; i++) Q[i * m + i] = T(1);
    const T small = std::is_same<T, float>::value ? T(1e-5) : T(1e-12);
    int end = m - 1, start = 0, iter = 0;
    while (end > 0) {
        for (int i = start; i < end; i++) {
            T a = std::abs(e[i]), b = std::abs(d[i]) + std::abs(d[i + 1]);
            if (a * a <= b * b * small * small) e[i] = T(0);
        }
        while (end > 0 && e[end - 1] == T(0)) end--;
        if (end <= 0) break;
        if (++iter > 30 * m) return false;
        start = end - 1;
        while (start > 0 && e[start - 1] != T(0)) start--;
        // one implicit QR sweep on rows/cols start..end
        T td = (d[end - 1] - d[end]) * T(0.5);
        T ee = e[end - 1];
        T mu = d[end];
        if (td == T(0)) mu -= std::abs(ee);
        else {
            T e2 = ee * ee;
            T h = std::sqrt(td * td + ee * ee);
            if (e2 == T(0)) mu -= (ee / (td + (td > T(0) ? T(1) : T(-1)))) * (ee / h);
            else mu -= e2 / (td + (td > T(0) ? h : -h));
        }
        T x = d[start] - mu, z = e[start];
        for (int k = start; k < end; ++k) {
            GivensPair<T> r = make_givens(x, z);
            const T c = r.c, s = r.s;
            T sdk = s * d[k] + c * e[k];
            T dkp1 = s * e[k] + c * d[k + 1];
            d[k] = c * (c * d[k] - s * e[k]) - s * (c * e[k] - s * d[k + 1]);
            d[k + 1] = s * sdk + c * dkp1;
            e[k] = c * sdk - s * dkp1;
            if (k > start) e[k - 1] = c * e[k - 1] - s * z;
            x = e[k];
            if (k < end - 1) { z = -s * e[k + 1]; e[k + 1] = c * e[k + 1]; }
            for (int i = 0; i < m; i++) {   // Q <- Q G on columns k, k+1
                T qi = Q[k * m + i], qj = Q[(k + 1) * m + i];
                Q[k * m + i] = c * qi - s * qj;
                Q[(k + 1) * m + i] = s * qi + c * qj;
            }
        }
    }
    return true;
}

struct LanczosInfo {
    int nmatvec = 0;      // operator applications
    int nrestart = 0;     // implicit restarts performed
    int converged = 0;    // 1 if the tol test passed (the reference reads garbage otherwise)
};

// op(v, w): w = S v for the symmetric n x n operator.  Returns the leading Ritz value.
// Throws nothing; returns NaN and sets info->converged = -1 for n < 3 (the reference's
// constructor would throw std::invalid_argument there).
template <class T, class Op>
T coarse_largest_eigenvalue(Op&& op, i64 n, LanczosInfo* info, int max_restart = 10, T tol = T(0.1), int ncv = 3)
{
    // ncv = 3 is what the reference's source has today; other values (2..8) exist for the README forensics in
    // tests/test_oracle_golden.py only (the README was knitted by a build with ncv = 2, see there).
    constexpr int MM = 8;
    const int m = std::min(std::max(ncv, 2), MM);
    const int nev = 1;
    LanczosInfo li;
    if (n < m) { li.converged = -1; if (info) *info = li; return std::numeric_limits<T>::quiet_NaN(); }
    const T prec = std::pow(std::numeric_limits<T>::epsilon(), T(2) / T(3));

    std::vector<T> V((size_t)n * m, T(0)), f(n), w(n), tmp(n);
    T H[MM * MM];                        // column-major; H(r,c) = H[c*m + r]
    for (int i = 0; i < m * m; i++) H[i] = T(0);
    auto Hat = [&](int r, int c) -> T& { return H[c * m + r]; };
    T ritz_val[MM] = {0}, ritz_est[MM] = {0};

    // --- start vector and first step (init) ---
    {
        LehmerStream rng(0);
        rng.fill(tmp.data(), n);
        T nv = norm2(tmp.data(), n);
        T* v0 = V.data();
        for (i64 i = 0; i < n; i++) v0[i] = tmp[i] / nv;
        op(v0, w.data()); li.nmatvec++;
        Hat(0, 0) = dot(v0, w.data(), n);
        for (i64 i = 0; i < n; i++) f[i] = w[i] - v0[i] * Hat(0, 0);
    }

    // extend the factorisation from column `from` to m columns, starting residual fk
    auto extend = [&](int from, const std::vector<T>& fk) {
        if (m <= from) return;
        f = fk;
        T beta = norm2(f.data(), n);
        for (int c = from; c < m; c++) for (int r = 0; r < m; r++) Hat(r, c) = T(0);
        for (int r = from; r < m; r++) for (int c = 0; c < from; c++) Hat(r, c) = T(0);
        for (int i = from; i <= m - 1; i++) {
            bool fresh = false;
            if (beta < prec) {
                // breakdown: draw a new direction orthogonal to the current basis
                LehmerStream rng(2 * i);
                rng.fill(f.data(), n);
                T Vf[MM];
                for (int c = 0; c < i; c++) Vf[c] = dot(V.data() + (size_t)c * n, f.data(), n);
                for (int c = 0; c < i; c++) {
                    const T* vc = V.data() + (size_t)c * n;
                    for (i64 r = 0; r < n; r++) f[r] -= vc[r] * Vf[c];
                }
                beta = norm2(f.data(), n);
                fresh = true;
            }
            T* v = V.data() + (size_t)i * n;
            for (i64 r = 0; r < n; r++) v[r] = f[r] / beta;
            Hat(i, i - 1) = fresh ? T(0) : beta;
            op(v, w.data()); li.nmatvec++;
            T Hii = dot(v, w.data(), n);
            Hat(i - 1, i) = Hat(i, i - 1);
            Hat(i, i) = Hii;
            const T* vp = V.data() + (size_t)(i - 1) * n;
            if (fresh) for (i64 r = 0; r < n; r++) f[r] = w[r] - Hii * v[r];
            else {
                const T hb = Hat(i, i - 1);
                for (i64 r = 0; r < n; r++) f[r] = w[r] - hb * vp[r] - Hii * v[r];
            }
            beta = norm2(f.data(), n);
            // re-orthogonalise against the first i+1 basis vectors, at most 5 passes
            T Vf[MM];
            auto project = [&]() {
                T mx = T(0);
                for (int c = 0; c <= i; c++) {
                    Vf[c] = dot(V.data() + (size_t)c * n, f.data(), n);
                    mx = std::max(mx, std::abs(Vf[c]));
                }
                return mx;
            };
            T mx = project();
            int count = 0;
            while (count < 5 && mx > prec * beta) {
                for (int c = 0; c <= i; c++) {
                    const T* vc = V.data() + (size_t)c * n;
                    for (i64 r = 0; r < n; r++) f[r] -= vc[r] * Vf[c];
                }
                Hat(i - 1, i) += Vf[i - 1];
                Hat(i, i - 1) = Hat(i - 1, i);
                Hat(i, i) += Vf[i];
                beta = norm2(f.data(), n);
                mx = project();
                count++;
            }
        }
    };

    // Ritz values of H sorted descending, with the last components of their vectors
    auto ritz = [&]() {
        T d[MM], e[MM], Q[MM * MM];
        for (int i = 0; i < m; i++) d[i] = Hat(i, i);
        for (int i = 0; i < m - 1; i++) e[i] = Hat(i + 1, i);
        tridiag_eig_small<T>(m, d, e, Q);
        int idx[MM];
        for (int i = 0; i < m; i++) idx[i] = i;
        for (int i = 1; i < m; i++) {                                   // descending (LARGEST_ALGE), insertion sort
            const int t = idx[i];
            int j = i;
            while (j > 0 && -d[t] < -d[idx[j - 1]]) { idx[j] = idx[j - 1]; j--; }
            idx[j] = t;
        }
        for (int i = 0; i < m; i++) { ritz_val[i] = d[idx[i]]; ritz_est[i] = Q[idx[i] * m + (m - 1)]; }
    };

    extend(1, f);
    ritz();

    int it, nconv = 0;
    for (it = 0; it < max_restart; it++) {
        // convergence of the wanted Ritz value
        T thresh = tol * std::max(std::abs(ritz_val[0]), prec);
        T resid = std::abs(ritz_est[0]) * norm2(f.data(), n);
        nconv = (resid < thresh) ? 1 : 0;
        if (nconv >= nev) break;
        // how many Ritz vectors to keep
        int k = nev;
        for (int i = nev; i < m; i++) if (std::abs(ritz_est[i]) < prec) k++;
        k += std::min(nconv, (m - k) / 2);
        if (k == 1 && m >= 6) k = m / 2; else if (k == 1 && m > 2) k = 2;
        if (k >= m) continue;            // nothing to restart with (reference: restart() returns)
        li.nrestart++;

        // shifted QR sweeps with the unwanted Ritz values
        T Q[MM * MM];
        for (int i = 0; i < m * m; i++) Q[i] = T(0);
        for (int i = 0; i < m; i++) Q[i * m + i] = T(1);
        for (int sidx = k; sidx < m; sidx++) {
            const T mu = ritz_val[sidx];
            for (int i = 0; i < m; i++) Hat(i, i) -= mu;
            // QR of the tridiagonal: Givens sequence (cs, sn), R kept in Tm
            T Tm[MM * MM];
            for (int i = 0; i < m * m; i++) Tm[i] = T(0);
            auto Tat = [&](int r, int c) -> T& { return Tm[c * m + r]; };
            for (int i = 0; i < m; i++) Tat(i, i) = Hat(i, i);
            for (int i = 0; i < m - 1; i++) { Tat(i, i + 1) = Hat(i + 1, i); Tat(i + 1, i) = Hat(i + 1, i); }
            T cs[MM], sn[MM];
            const T eps = std::numeric_limits<T>::epsilon();
            for (int i = 0; i < m - 1; i++) {
                T a = Tat(i, i), b = Tat(i + 1, i);
                T r = std::sqrt(a * a + b * b);
                if (r <= eps) { r = T(0); cs[i] = T(1); sn[i] = T(0); }
                else { cs[i] = a / r; sn[i] = -b / r; }
                Tat(i, i) = r; Tat(i + 1, i) = T(0);
                T t = Tat(i, i + 1);
                Tat(i, i + 1) = cs[i] * t - sn[i] * Tat(i + 1, i + 1);
                Tat(i + 1, i + 1) = sn[i] * t + cs[i] * Tat(i + 1, i + 1);
                if (i < m - 2) {
                    Tat(i, i + 2) = -sn[i] * Tat(i + 1, i + 2);
                    Tat(i + 1, i + 2) *= cs[i];
                }
            }
            // Q <- Q * G_0 * G_1 ...
            for (int i = 0; i < m - 1; i++)
                for (int r = 0; r < m; r++) {
                    T t = Q[i * m + r];
                    Q[i * m + r] = cs[i] * t - sn[i] * Q[(i + 1) * m + r];
                    Q[(i + 1) * m + r] = sn[i] * t + cs[i] * Q[(i + 1) * m + r];
                }
            // H <- R Q (tridiagonal again), then undo the shift
            T RQ[MM * MM];
            for (int i = 0; i < m * m; i++) RQ[i] = T(0);
            auto Rat = [&](int r, int c) -> T& { return RQ[c * m + r]; };
            for (int i = 0; i < m; i++) Rat(i, i) = Tat(i, i);
            for (int i = 0; i < m - 1; i++) Rat(i, i + 1) = Tat(i, i + 1);
            for (int i = 0; i < m - 1; i++) {
                T m11 = Rat(i, i), m12 = Rat(i, i + 1), m21 = Rat(i + 1, i), m22 = Rat(i + 1, i + 1);
                Rat(i, i) = cs[i] * m11 - sn[i] * m12;
                Rat(i + 1, i) = cs[i] * m21 - sn[i] * m22;
                Rat(i + 1, i + 1) = sn[i] * m21 + cs[i] * m22;
            }
            for (int i = 0; i < m - 1; i++) Rat(i, i + 1) = Rat(i + 1, i);
            for (int i = 0; i < m * m; i++) H[i] = RQ[i];
            for (int i = 0; i < m; i++) Hat(i, i) += mu;
        }
        // V <- V Q for the leading k+1 columns (column i of Q has m-k+i+1 leading non-zeros)
        std::vector<T> Vs((size_t)n * (k + 1), T(0));
        for (int i = 0; i <= k; i++) {
            const int nnz = (i < k) ? (m - k + i + 1) : m;
            T* dst = Vs.data() + (size_t)i * n;
            for (int c = 0; c < nnz; c++) {
                const T q = Q[i * m + c];
                const T* vc = V.data() + (size_t)c * n;
                for (i64 r = 0; r < n; r++) dst[r] += vc[r] * q;
            }
        }
        std::copy(Vs.begin(), Vs.end(), V.begin());
        std::vector<T> fk(n);
        const T qf = Q[(k - 1) * m + (m - 1)], hk = Hat(k, k - 1);
        const T* vk = V.data() + (size_t)k * n;
        for (i64 r = 0; r < n; r++) fk[r] = f[r] * qf + vk[r] * hk;
        extend(k, fk);
        ritz();
    }
    li.converged = nconv >= nev ? 1 : 0;
    if (info) *info = li;
    return ritz_val[0];
}

}  // namespace oracle
