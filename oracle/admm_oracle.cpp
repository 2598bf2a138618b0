// oracle/admm_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's (yixuan/ADMM) solvers, used ONLY as the parity
// checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference leg.  Nothing under admm_b200/ may include, link or call this file.
//
// The reference itself cannot be compiled in this image (needs R, Rcpp, RcppEigen and
// Eigen headers, none present), so this is a restatement from reading
//   /root/reference/src/FADMMBase.h:185-265      accelerated ADMM loop
//   /root/reference/src/ADMMBase.h:73-109,158-216 plain ADMM loop, rho balancing
//   /root/reference/src/ADMMLassoTall.h:55-230    lasso n > p (float)
//   /root/reference/src/ADMMLassoWide.h:70-251    lasso n <= p (float, linearised)
//   /root/reference/src/ADMMEnet.h:19-154         elastic-net prox swaps
//   /root/reference/src/ADMMLAD.h:62-225          least absolute deviation (double)
//   /root/reference/src/ADMMBP.h:48-197           basis pursuit (double)
//   /root/reference/src/PADMMBase.h:57-237, PADMMLasso.h:17-223  row-split consensus lasso (float)
//   /root/reference/src/DataStd.h:39-207          standardisation / recovery
//   /root/reference/src/Lasso.cpp:39-135, Enet.cpp, ParLasso.cpp, LAD.cpp, BP.cpp   entry points
// Parity pinning: tests/test_oracle_golden.py checks this file against the five result
// vectors printed in /root/reference/README.md; tests/test_oracle_readme_benchmarks.py
// against the coefficient-difference ranges the README prints for its timing sections
// (BP / LAD: every printed digit; parallel lasso: 7 digits; the wide lasso / enet rows
// of README.md:287-289 to 1e-6 / one float32 ulp, with glmnet restated at its default
// threshold in tests/glmnet_naive.py).  These are the only reference-produced outputs
// that exist; fixtures and the generating scripts live in tests/golden/.
//
// Mixed precision follows the reference: scalars (rho, eps, residuals, acceleration
// coefficients) are double; lasso/enet/consensus vectors are float; LAD/BP are double.
// z-type vectors are kept dense here -- the reference's SparseVector bookkeeping changes
// no arithmetic result because its sparse reductions add the same terms in the same
// (index) order and adding an exact zero is exact; support := {i : z_i != 0}.
#include "linalg.hpp"
#include "lanczos.hpp"
#include "synth_stream.hpp"
#include <cstdio>
#include <chrono>

using namespace oracle;

namespace {

// ncv of the Spectra call behind the default rho / gamma.  3 = the reference's source as it stands (ADMMLassoTall.h:196,
// ADMMLassoWide.h:202); settable through oracle_set_lanczos_ncv for the README forensics of tests/test_oracle_golden.py only.
int g_lanczos_ncv = 3;

struct Trace {           // optional per-iteration record for parity tests
    double* buf = nullptr;   // rows of 5: eps_primal, resid_primal, eps_dual, resid_dual, rho
    int cap = 0;
    int n = 0;
    void push(double ep, double rp, double ed, double rd, double rho)
    {
        if (buf && n < cap) {
            double* r = buf + 5 * (size_t)n;
            r[0] = ep; r[1] = rp; r[2] = ed; r[3] = rd; r[4] = rho;
        }
        n++;
    }
};

// ---------------------------------------------------------------------------------
// DataStd  (/root/reference/src/DataStd.h:89-207, non-AVX branch)
// ---------------------------------------------------------------------------------
template <class T> struct Standardizer {
    int flag; i64 n, p;
    T meanY = 0, scaleY = 1;
    std::vector<T> meanX, scaleX;
    Standardizer(i64 n_, i64 p_, bool standardize, bool intercept)
        : flag(int(standardize) + 2 * int(intercept)), n(n_), p(p_)
    {
        if (flag == 3 || flag == 2) meanX.assign(p, T(0));
        if (flag == 3 || flag == 1) scaleX.assign(p, T(1));
    }
    static T sd_n(const T* v, i64 len)
    {
        const T mu = mean(v, len);
        T s0 = 0, s1 = 0, s2 = 0, s3 = 0; i64 i = 0;
        for (; i + 4 <= len; i += 4) {
            T a = v[i] - mu, b = v[i + 1] - mu, c = v[i + 2] - mu, d = v[i + 3] - mu;
            s0 += a * a; s1 += b * b; s2 += c * c; s3 += d * d;
        }
        for (; i < len; i++) { T a = v[i] - mu; s0 += a * a; }
        return std::sqrt((s0 + s1) + (s2 + s3)) / std::sqrt(T(len));
    }
    void apply(T* X, T* Y)
    {
        const T n_invsqrt = T(1.0 / std::sqrt(T(n)));
        switch (flag) {
        case 1:
            scaleY = sd_n(Y, n);
            for (i64 i = 0; i < n; i++) Y[i] /= scaleY;
            break;
        case 2: case 3:
            meanY = mean(Y, n);
            for (i64 i = 0; i < n; i++) Y[i] -= meanY;
            scaleY = norm2(Y, n) * n_invsqrt;
            for (i64 i = 0; i < n; i++) Y[i] /= scaleY;
            break;
        default: break;
        }
        if (flag == 0) return;
#pragma omp parallel for schedule(static) if (n * p > 100000)
        for (i64 j = 0; j < p; j++) {
            T* c = X + j * n;
            if (flag == 1) {
                scaleX[j] = sd_n(c, n);
                const T inv = T(1.0 / scaleX[j]);
                for (i64 i = 0; i < n; i++) c[i] *= inv;
            } else if (flag == 2) {
                meanX[j] = mean(c, n);
                const T mu = meanX[j];
                for (i64 i = 0; i < n; i++) c[i] -= mu;
            } else {
                meanX[j] = mean(c, n);
                const T mu = meanX[j];
                for (i64 i = 0; i < n; i++) c[i] -= mu;
                scaleX[j] = norm2(c, n) * n_invsqrt;
                const T inv = T(1.0 / scaleX[j]);
                for (i64 i = 0; i < n; i++) c[i] *= inv;
            }
        }
    }
    // coef (length p, zeros = not in support) -> original scale; returns intercept
    T recover(T* coef) const
    {
        T beta0 = 0;
        switch (flag) {
        case 1:
            for (i64 j = 0; j < p; j++) if (coef[j] != T(0)) { coef[j] /= scaleX[j]; coef[j] *= scaleY; }
            break;
        case 2: {
            T s = 0;
            for (i64 j = 0; j < p; j++) if (coef[j] != T(0)) { coef[j] *= scaleY; s += coef[j] * meanX[j]; }
            beta0 = meanY - s;
        } break;
        case 3: {
            T s = 0;
            for (i64 j = 0; j < p; j++) if (coef[j] != T(0)) { coef[j] /= scaleX[j]; coef[j] *= scaleY; s += coef[j] * meanX[j]; }
            beta0 = meanY - s;
        } break;
        default: break;
        }
        return beta0;
    }
};

// rho balancing shared by both loops (/root/reference/src/FADMMBase.h:109-133)
inline void balance_rho(double& rho, double rp, double ep, double rd, double ed)
{
    if (rp / ep > 10 * rd / ed) rho *= 2;
    else if (rd / ed > 10 * rp / ep) rho /= 2;
    if (rp < ep) rho /= 1.2;
    if (rd < ed) rho *= 1.2;
}

// ---------------------------------------------------------------------------------
// Accelerated loop  (/root/reference/src/FADMMBase.h:219-265).  Model supplies the hooks.
// State vectors live in the model; V is the dual/aux scalar type.
// ---------------------------------------------------------------------------------
template <class V> struct FastState {
    i64 dz = 0, dy = 0;
    std::vector<V> z, y, adj_z, adj_y, old_z, old_y;
    double adj_a = 1.0, adj_c = 9999;
    double rho = 1, eps_primal = 0, eps_dual = 0, resid_primal = 9999, resid_dual = 9999;
    void alloc(i64 dz_, i64 dy_)
    {
        dz = dz_; dy = dy_;
        z.assign(dz, 0); adj_z.assign(dz, 0); old_z.assign(dz, 0);
        y.assign(dy, 0); adj_y.assign(dy, 0); old_y.assign(dy, 0);
    }
};

template <class V, class Model> int fast_admm_solve(Model& m, FastState<V>& s, int maxit, bool rho_adapts, Trace* tr)
{
    int i;
    for (i = 0; i < maxit; i++) {
        s.old_z = s.z;
        s.old_y = s.y;
        // update_x: tolerances from the iterate *before* this step
        s.eps_primal = m.eps_primal();
        s.eps_dual = m.eps_dual();
        m.step_x();
        // update_z
        m.step_z();
        s.resid_dual = s.rho * std::sqrt((double)diff_sqnorm_seq(s.z.data(), s.old_z.data(), s.dz));
        // update_y
        s.resid_primal = m.step_residual_and_dual();
        if (tr) tr->push(s.eps_primal, s.resid_primal, s.eps_dual, s.resid_dual, s.rho);
        if (s.resid_primal < s.eps_primal && s.resid_dual < s.eps_dual) break;

        const double old_c = s.adj_c;
        s.adj_c = s.rho * s.resid_primal * s.resid_primal +
                  s.rho * (double)diff_sqnorm_seq(s.z.data(), s.adj_z.data(), s.dz);
        if (s.adj_c < 0.999 * old_c) {
            const double old_a = s.adj_a;
            s.adj_a = 0.5 + 0.5 * std::sqrt(1 + 4.0 * old_a * old_a);
            const double ratio = (old_a - 1.0) / s.adj_a;
            const V c1 = V(1 + ratio), c2 = V(ratio);
            for (i64 k = 0; k < s.dz; k++) s.adj_z[k] = c1 * s.z[k] - c2 * s.old_z[k];
            for (i64 k = 0; k < s.dy; k++) s.adj_y[k] = c1 * s.y[k] - c2 * s.old_y[k];
        } else {
            s.adj_a = 1.0;
            s.adj_z = s.old_z;
            s.adj_y = s.old_y;
            s.adj_c = old_c / 0.999;
        }
        if (i > 5 && rho_adapts) balance_rho(s.rho, s.resid_primal, s.eps_primal, s.resid_dual, s.eps_dual);
    }
    return i + 1;
}

// ---------------------------------------------------------------------------------
// Lasso / elastic net, n > p   (/root/reference/src/ADMMLassoTall.h, ADMMEnet.h:19-58)
// Works from the Gram matrix: after setup the reference never touches X again.
// ---------------------------------------------------------------------------------
struct TallLasso {
    i64 p;
    const float* XY;          // X'y
    std::vector<float> L;     // factor of X'X + rho I (lower, p x p)
    std::vector<float> x, rhs, r;
    FastState<float> s;
    float lambda = 0, lambda0 = 0;
    bool enet = false; float alpha = 1;
    double eps_abs, eps_rel;
    double ev_estimate = 0; LanczosInfo lz;

    TallLasso(i64 p_, const float* XY_, double ea, double er) : p(p_), XY(XY_), eps_abs(ea), eps_rel(er)
    {
        float m = 0;
        for (i64 i = 0; i < p; i++) m = std::max(m, std::abs(XY[i]));
        lambda0 = m;
    }
    void set_enet(double alpha_) { enet = true; alpha = (float)alpha_; lambda0 = float(lambda0 / (alpha + 0.0001)); }

    // G: lower(X'X), p x p column-major; consumed (overwritten by the factor)
    int init(std::vector<float>& G, double lambda_, double rho_)
    {
        x.assign(p, 0.f); rhs.assign(p, 0.f); r.assign(p, 0.f);
        s.alloc(p, p);
        lambda = (float)lambda_;
        s.rho = rho_;
        if (s.rho <= 0) {
            const float* Gp = G.data(); const i64 pp = p;
            float ev = coarse_largest_eigenvalue<float>(
                [Gp, pp](const float* v, float* w) { symv_lower(Gp, pp, v, w); }, p, &lz, 10, 0.1f, g_lanczos_ncv);
            if (lz.converged < 0) return -3;
            ev_estimate = ev;
            s.rho = std::pow((double)ev, 1.0 / 3) * std::pow((double)lambda, 2.0 / 3);
        }
        for (i64 i = 0; i < p; i++) G[i * p + i] += (float)s.rho;   // float += double -> rounded once
        L.swap(G);
        int info = chol_lower(L.data(), p);
        s.eps_primal = 0; s.eps_dual = 0; s.resid_primal = 9999; s.resid_dual = 9999;
        s.adj_a = 1.0; s.adj_c = 9999;
        return info;
    }
    void init_warm(double lambda_)
    {
        lambda = (float)lambda_;
        s.eps_primal = 0; s.eps_dual = 0; s.resid_primal = 9999; s.resid_dual = 9999;
        // adj_a / adj_c deliberately carried over (ADMMLassoTall.h:228-229)
    }
    double eps_primal() const
    {
        double rr = std::max(norm2(x.data(), p), std::sqrt(sqnorm_seq(s.z.data(), p)));
        return rr * eps_rel + std::sqrt(double(p)) * eps_abs;
    }
    double eps_dual() const { return norm2(s.y.data(), p) * eps_rel + std::sqrt(double(p)) * eps_abs; }
    void step_x()
    {
        for (i64 i = 0; i < p; i++) {
            float v = XY[i] - s.adj_y[i];
            if (s.adj_z[i] != 0.f) v = float(double(v) + s.rho * double(s.adj_z[i]));
            rhs[i] = v;
        }
        chol_solve(L.data(), p, rhs.data());
        x = rhs;
    }
    void step_z()
    {
        const float frho = (float)s.rho;
        const double pen = double(lambda) / s.rho;
        if (!enet) {
            for (i64 i = 0; i < p; i++) {
                const float v = x[i] + s.adj_y[i] / frho;
                if (v > pen) s.z[i] = float(v - pen);
                else if (v < -pen) s.z[i] = float(v + pen);
                else s.z[i] = 0.f;
            }
        } else {
            const float thresh = float(alpha * pen);
            const float denom = float(1.0 + pen * (1.0 - alpha));
            for (i64 i = 0; i < p; i++) {
                const float v = x[i] + s.adj_y[i] / frho;
                if (v > thresh) s.z[i] = (v - thresh) / denom;
                else if (v < -thresh) s.z[i] = (v + thresh) / denom;
                else s.z[i] = 0.f;
            }
        }
    }
    double step_residual_and_dual()
    {
        for (i64 i = 0; i < p; i++) r[i] = x[i] - s.z[i];
        const double rp = norm2(r.data(), p);
        const float frho = (float)s.rho;
        for (i64 i = 0; i < p; i++) s.y[i] = s.adj_y[i] + frho * r[i];
        return rp;
    }
    int solve(int maxit, Trace* tr) { return fast_admm_solve<float>(*this, s, maxit, /*rho_adapts=*/false, tr); }
};

// ---------------------------------------------------------------------------------
// Lasso / elastic net, n <= p   (/root/reference/src/ADMMLassoWide.h, ADMMEnet.h:62-154)
// ---------------------------------------------------------------------------------
struct WideLasso {
    i64 n, p;
    const float* X; const float* Y;
    std::vector<float> x;             // length p, zeros = not in support
    std::vector<i64> supp;            // sorted support of x
    std::vector<float> Ax, z, y, tmp, newz, r, vec;
    float sprad = 0, lambda = 0, lambda0 = 0;
    bool enet = false; float alpha = 1;
    int iter_counter = 0;
    double rho = 1, eps_abs, eps_rel, eps_p = 0, eps_d = 0, res_p = 9999, res_d = 9999;
    LanczosInfo lz;

    WideLasso(const float* X_, const float* Y_, i64 n_, i64 p_, double ea, double er)
        : n(n_), p(p_), X(X_), Y(Y_), eps_abs(ea), eps_rel(er)
    {
        vec.resize(p);
        gemv_t(X, n, p, Y, vec.data());
        float m = 0;
        for (i64 i = 0; i < p; i++) m = std::max(m, std::abs(vec[i]));
        lambda0 = m;
        std::vector<float> G((size_t)n * n);
        gram_nt_lower(X, n, p, G.data());
        const float* Gp = G.data(); const i64 nn = n;
        sprad = coarse_largest_eigenvalue<float>(
            [Gp, nn](const float* v, float* w) { symv_lower(Gp, nn, v, w); }, n, &lz, 10, 0.1f, g_lanczos_ncv);
    }
    void set_enet(double alpha_) { enet = true; alpha = (float)alpha_; lambda0 = float(lambda0 / (alpha + 0.0001)); }
    void init(double lambda_, double rho_)
    {
        x.assign(p, 0.f); supp.clear();
        Ax.assign(n, 0.f); z.assign(n, 0.f); y.assign(n, 0.f);
        tmp.assign(n, 0.f); newz.assign(n, 0.f); r.assign(n, 0.f);
        lambda = (float)lambda_;
        rho = rho_;
        if (rho <= 0) rho = std::pow(double(lambda / sprad), 1.0 / 3);
        eps_p = eps_d = 0; res_p = res_d = 9999;
        iter_counter = 0;
    }
    void init_warm(double lambda_)
    {
        lambda = (float)lambda_;
        eps_p = eps_d = 0; res_p = res_d = 9999;
        iter_counter = 0;
    }
    static bool is_regular(unsigned c)
    {
        if (c == 0 || c == 3 || c == 15 || c == 63) return true;
        c++;
        if (c & (c - 1)) return false;
        return (c & 0x55555555u) != 0;
    }
    // prox on one value with float threshold semantics (active-set branch)
    void rebuild_support()
    {
        supp.clear();
        for (i64 j = 0; j < p; j++) if (x[j] != 0.f) supp.push_back(j);
    }
    void step_x()
    {
        if (!enet && double(lambda) > double(lambda0) - 1e-5) {
            for (i64 j : supp) x[j] = 0.f;
            supp.clear();
            return;                                  // counter not advanced (ADMMLassoWide.h:131-135)
        }
        const float gamma = sprad;
        const float frho = (float)rho;
        const bool regular = is_regular((unsigned)iter_counter) && (!enet || lambda < lambda0);
        if (regular) {
            for (i64 i = 0; i < n; i++) tmp[i] = Ax[i] + z[i] + y[i] / frho;
            gemv_t(X, n, p, tmp.data(), vec.data());
            const double pen = double(lambda) / (rho * double(gamma));
            if (!enet) {
                for (i64 j = 0; j < p; j++) {
                    const float v = -vec[j] / gamma + x[j];
                    if (v > pen) x[j] = float(v - pen);
                    else if (v < -pen) x[j] = float(v + pen);
                    else x[j] = 0.f;
                }
            } else {
                const float thresh = float(alpha * pen);
                const float denom = float(1.0 + pen * (1.0 - alpha));
                for (i64 j = 0; j < p; j++) {
                    const float v = -vec[j] / gamma + x[j];
                    if (v > thresh) x[j] = (v - thresh) / denom;
                    else if (v < -thresh) x[j] = (v + thresh) / denom;
                    else x[j] = 0.f;
                }
            }
            rebuild_support();
        } else {
            const float pen = float(double(lambda) / (rho * double(gamma)));
            const float thresh = float(alpha * pen);
            const float denom = float(1.0 + pen * (1.0 - alpha));
            for (i64 i = 0; i < n; i++) tmp[i] = (Ax[i] + z[i] + y[i] / frho) / gamma;
            const i64 nnz = (i64)supp.size();
#pragma omp parallel for schedule(static) if (nnz * n > 100000)
            for (i64 k = 0; k < nnz; k++) {
                const i64 j = supp[k];
                const float v = x[j] - dot(tmp.data(), X + j * n, n);
                if (!enet) {
                    if (v > pen) x[j] = v - pen;
                    else if (v < -pen) x[j] = v + pen;
                    else x[j] = 0.f;
                } else {
                    if (v > thresh) x[j] = (v - thresh) / denom;
                    else if (v < -thresh) x[j] = (v + thresh) / denom;
                    else x[j] = 0.f;
                }
            }
            std::vector<i64> keep; keep.reserve(supp.size());
            for (i64 j : supp) if (x[j] != 0.f) keep.push_back(j);
            supp.swap(keep);
        }
        iter_counter++;
    }
    int solve(int maxit, Trace* tr)
    {
        int i;
        for (i = 0; i < maxit; i++) {
            // tolerances from the previous iterate (stale Ax on purpose)
            eps_p = std::max(norm2(Ax.data(), n), norm2(z.data(), n)) * eps_rel + std::sqrt(double(n)) * eps_abs;
            eps_d = double(std::sqrt(sprad) * norm2(y.data(), n)) * eps_rel + std::sqrt(double(p)) * eps_abs;
            step_x();
            // z-step
            std::fill(Ax.begin(), Ax.end(), 0.f);
            for (i64 j : supp) {
                const float v = x[j]; const float* c = X + j * n;
                for (i64 k = 0; k < n; k++) Ax[k] += c[k] * v;
            }
            const float frho = (float)rho, den = float(-1 - rho);
            for (i64 k = 0; k < n; k++) newz[k] = (Y[k] + y[k] + frho * Ax[k]) / den;
            {
                float sdiff = 0;   // (new_z - z).norm()
                float s0 = 0, s1 = 0, s2 = 0, s3 = 0; i64 k = 0;
                for (; k + 4 <= n; k += 4) {
                    float a = newz[k] - z[k], b = newz[k + 1] - z[k + 1], c = newz[k + 2] - z[k + 2], d = newz[k + 3] - z[k + 3];
                    s0 += a * a; s1 += b * b; s2 += c * c; s3 += d * d;
                }
                for (; k < n; k++) { float a = newz[k] - z[k]; s0 += a * a; }
                sdiff = std::sqrt((s0 + s1) + (s2 + s3));
                res_d = rho * double(std::sqrt(sprad)) * double(sdiff);
            }
            z.swap(newz);
            for (i64 k = 0; k < n; k++) r[k] = Ax[k] + z[k];
            res_p = norm2(r.data(), n);
            for (i64 k = 0; k < n; k++) y[k] += frho * r[k];
            if (tr) tr->push(eps_p, res_p, eps_d, res_d, rho);
            if (res_p < eps_p && res_d < eps_d) break;
            if (i > 3) balance_rho(rho, res_p, eps_p, res_d, eps_d);
        }
        return i + 1;
    }
};

// ---------------------------------------------------------------------------------
// Row-split consensus lasso  (/root/reference/src/PADMMBase.h, PADMMLasso.h)
// ---------------------------------------------------------------------------------
struct ConsensusLasso {
    struct Worker {
        i64 rows, p;
        std::vector<float> A, b, Ab, L, x, y, rhs, t1;
        bool tall;
        double sq_resid = 0;
    };
    i64 n, p; int N;
    std::vector<Worker> w;
    std::vector<float> z, newz, acc;
    double lambda = 0, lambda0 = 0, rho = 1, eps_abs, eps_rel;
    double eps_p = 0, eps_d = 0, res_p = 9999, res_d = 9999;

    ConsensusLasso(const float* X, const float* Y, i64 n_, i64 p_, int N_, double ea, double er)
        : n(n_), p(p_), N(N_), w(N_), eps_abs(ea), eps_rel(er)
    {
        std::vector<float> xy(p);
        gemv_t(X, n, p, Y, xy.data());
        float m = 0;
        for (i64 i = 0; i < p; i++) m = std::max(m, std::abs(xy[i]));
        lambda0 = m;
        const i64 chunk = n / N, last = chunk + n % N;
        for (int k = 0; k < N; k++) {
            Worker& wk = w[k];
            wk.rows = (k < N - 1) ? chunk : last; wk.p = p;
            wk.tall = wk.rows >= p;
            wk.A.resize((size_t)wk.rows * p); wk.b.resize(wk.rows);
            for (i64 j = 0; j < p; j++)
                std::copy(X + j * n + k * chunk, X + j * n + k * chunk + wk.rows, wk.A.data() + j * wk.rows);
            std::copy(Y + k * chunk, Y + k * chunk + wk.rows, wk.b.data());
            wk.Ab.resize(p);
            gemv_t(wk.A.data(), wk.rows, p, wk.b.data(), wk.Ab.data());
            wk.x.assign(p, 0.f); wk.y.assign(p, 0.f); wk.rhs.assign(p, 0.f); wk.t1.assign(wk.rows, 0.f);
        }
        z.assign(p, 0.f); newz.assign(p, 0.f); acc.assign(p, 0.f);
    }
    int init(double lambda_, double rho_)
    {
        std::fill(z.begin(), z.end(), 0.f);
        lambda = lambda_;
        rho = rho_;
        if (rho <= 0) rho = lambda / N;
        int bad = 0;
        for (auto& wk : w) {
            std::fill(wk.x.begin(), wk.x.end(), 0.f);
            std::fill(wk.y.begin(), wk.y.end(), 0.f);
            const i64 d = wk.tall ? p : wk.rows;
            wk.L.assign((size_t)d * d, 0.f);
            if (wk.tall) gram_tn_lower(wk.A.data(), wk.rows, p, wk.L.data());
            else gram_nt_lower(wk.A.data(), wk.rows, p, wk.L.data());
            for (i64 i = 0; i < d; i++) wk.L[i * d + i] += (float)rho;
            if (chol_lower(wk.L.data(), d) != 0) bad = 1;
        }
        eps_p = eps_d = 0; res_p = res_d = 9999;
        return bad;
    }
    void init_warm(double lambda_) { lambda = lambda_; eps_p = eps_d = 0; res_p = res_d = 9999; }
    int solve(int maxit, Trace* tr)
    {
        int it;
        for (it = 0; it < maxit; it++) {
            double xs = 0, ys = 0;
            for (auto& wk : w) { xs += (double)sqnorm(wk.x.data(), p); ys += (double)sqnorm(wk.y.data(), p); }
            const double znorm = std::sqrt(sqnorm_seq(z.data(), p));
            eps_p = std::max(std::sqrt(xs), znorm * std::sqrt((double)N)) * eps_rel + std::sqrt(double(p * N)) * eps_abs;
            eps_d = std::sqrt(ys) * eps_rel + std::sqrt(double(p * N)) * eps_abs;
#pragma omp parallel for schedule(static)
            for (int k = 0; k < N; k++) {
                Worker& wk = w[k];
                for (i64 i = 0; i < p; i++) {
                    float v = wk.Ab[i] - wk.y[i];
                    if (z[i] != 0.f) v = float(double(v) + rho * double(z[i]));
                    wk.rhs[i] = v;
                }
                if (wk.tall) {
                    wk.x = wk.rhs;
                    chol_solve(wk.L.data(), p, wk.x.data());
                } else {
                    gemv_n(wk.A.data(), wk.rows, p, wk.rhs.data(), wk.t1.data());
                    chol_solve(wk.L.data(), wk.rows, wk.t1.data());
                    gemv_t(wk.A.data(), wk.rows, p, wk.t1.data(), wk.x.data());
                    const float frho = (float)rho;
                    for (i64 i = 0; i < p; i++) wk.x[i] = (wk.rhs[i] - wk.x[i]) / frho;
                }
            }
            // master z-update
            std::fill(acc.begin(), acc.end(), 0.f);
            const float frho = (float)rho;
            for (auto& wk : w) for (i64 i = 0; i < p; i++) acc[i] += wk.x[i] + wk.y[i] / frho;
            const float fN = (float)N;
            const double pen = lambda / (rho * N);
            for (i64 i = 0; i < p; i++) {
                const float v = acc[i] / fN;
                if (v > pen) newz[i] = float(v - pen);
                else if (v < -pen) newz[i] = float(v + pen);
                else newz[i] = 0.f;
            }
            res_d = rho * std::sqrt(N * (double)diff_sqnorm_seq(newz.data(), z.data(), p));
            z.swap(newz);
            double rsum = 0;
            for (auto& wk : w) {
                float s0 = 0, s1 = 0, s2 = 0, s3 = 0; i64 i = 0;
                for (; i + 4 <= p; i += 4) {
                    float a = wk.x[i] - z[i], b = wk.x[i + 1] - z[i + 1], c = wk.x[i + 2] - z[i + 2], d = wk.x[i + 3] - z[i + 3];
                    s0 += a * a; s1 += b * b; s2 += c * c; s3 += d * d;
                    wk.y[i] += frho * a; wk.y[i + 1] += frho * b; wk.y[i + 2] += frho * c; wk.y[i + 3] += frho * d;
                }
                for (; i < p; i++) { float a = wk.x[i] - z[i]; s0 += a * a; wk.y[i] += frho * a; }
                wk.sq_resid = (double)((s0 + s1) + (s2 + s3));
                rsum += wk.sq_resid;
            }
            res_p = std::sqrt(rsum);
            if (tr) tr->push(eps_p, res_p, eps_d, res_d, rho);
            if (res_p < eps_p && res_d < eps_d) break;
        }
        return it + 1;
    }
};

// ---------------------------------------------------------------------------------
// LAD  (/root/reference/src/ADMMLAD.h) -- double; variable is xx = X beta in R^n
// ---------------------------------------------------------------------------------
struct LadModel {
    i64 n, p;
    const double* X; const double* Y;
    std::vector<double> L, H, x, r, v, t;
    FastState<double> s;
    double ynorm, eps_abs, eps_rel;
    bool use_hat;
    LadModel(const double* X_, const double* Y_, i64 n_, i64 p_, double rho_, double ea, double er)
        : n(n_), p(p_), X(X_), Y(Y_), eps_abs(ea), eps_rel(er)
    {
        ynorm = norm2(Y, n);
        L.assign((size_t)p * p, 0.0);
        gram_tn_lower(X, n, p, L.data());
        chol_lower(L.data(), p);
        use_hat = n <= 2000;
        if (use_hat) {
            std::vector<double> T(X, X + (size_t)n * p);
            trsm_right_lower_trans(L.data(), p, T.data(), n);     // T = X L^{-T}
            H.assign((size_t)n * n, 0.0);
            gram_nt_lower(T.data(), n, p, H.data());              // H = T T'
        }
        x.assign(n, 0.0); r.assign(n, 0.0); v.assign(n, 0.0); t.assign(p, 0.0);
        s.alloc(n, n);
        s.rho = rho_;
    }
    double eps_primal() const
    {
        double rr = std::max(norm2(x.data(), n), std::sqrt(sqnorm_seq(s.z.data(), n)));
        rr = std::max(rr, ynorm);
        return rr * eps_rel + std::sqrt(double(n)) * eps_abs;
    }
    double eps_dual() const { return norm2(s.y.data(), n) * eps_rel + std::sqrt(double(n)) * eps_abs; }
    void step_x()
    {
        for (i64 i = 0; i < n; i++) v[i] = (Y[i] - s.adj_y[i] / s.rho) + s.adj_z[i];
        if (use_hat) symv_lower(H.data(), n, v.data(), x.data());
        else {
            gemv_t(X, n, p, v.data(), t.data());
            chol_solve(L.data(), p, t.data());
            gemv_n(X, n, p, t.data(), x.data());
        }
    }
    void step_z()
    {
        const double pen = 1.0 / s.rho;
        for (i64 i = 0; i < n; i++) {
            const double u = (x[i] - Y[i]) + s.adj_y[i] / s.rho;
            if (u > pen) s.z[i] = u - pen; else if (u < -pen) s.z[i] = u + pen; else s.z[i] = 0.0;
        }
    }
    double step_residual_and_dual()
    {
        for (i64 i = 0; i < n; i++) r[i] = (x[i] - Y[i]) - s.z[i];
        const double rp = norm2(r.data(), n);
        for (i64 i = 0; i < n; i++) s.y[i] = s.adj_y[i] + s.rho * r[i];
        return rp;
    }
    int solve(int maxit, Trace* tr) { return fast_admm_solve<double>(*this, s, maxit, true, tr); }
    void coefficients(double* beta)
    {
        for (i64 i = 0; i < n; i++) v[i] = (Y[i] - s.adj_y[i] / s.rho) + s.adj_z[i];
        gemv_t(X, n, p, v.data(), beta);
        chol_solve(L.data(), p, beta);
    }
};

// ---------------------------------------------------------------------------------
// Basis pursuit  (/root/reference/src/ADMMBP.h) -- double; A is n x p with p > n
// ---------------------------------------------------------------------------------
struct BpModel {
    i64 n, p;
    std::vector<double> M, q, x, r, v, wk;
    FastState<double> s;
    double eps_abs, eps_rel;
    BpModel(const double* A, const double* b, i64 n_, i64 p_, double rho_, double ea, double er)
        : n(n_), p(p_), eps_abs(ea), eps_rel(er)
    {
        std::vector<double> L((size_t)n * n, 0.0);
        gram_nt_lower(A, n, p, L.data());
        chol_lower(L.data(), n);
        std::vector<double> t(b, b + n);
        chol_solve(L.data(), n, t.data());
        q.assign(p, 0.0);
        gemv_t(A, n, p, t.data(), q.data());                    // A'(AA')^{-1} b
        M.assign(A, A + (size_t)n * p);
        trsm_left_lower_notrans(L.data(), n, M.data(), p);      // M = L^{-1} A
        x.assign(p, 0.0); r.assign(p, 0.0); v.assign(p, 0.0); wk.assign(n, 0.0);
        s.alloc(p, p);
        s.rho = rho_;
    }
    double eps_primal() const
    {
        double rr = std::max(norm2(x.data(), p), std::sqrt(sqnorm_seq(s.z.data(), p)));
        return rr * eps_rel + std::sqrt(double(p)) * eps_abs;
    }
    double eps_dual() const { return norm2(s.y.data(), p) * eps_rel + std::sqrt(double(p)) * eps_abs; }
    void step_x()
    {
        for (i64 i = 0; i < p; i++) { v[i] = -s.adj_y[i] / s.rho; if (s.adj_z[i] != 0.0) v[i] += s.adj_z[i]; }
        for (i64 i = 0; i < p; i++) x[i] = v[i] + q[i];
        gemv_n(M.data(), n, p, v.data(), wk.data());
        gemv_t(M.data(), n, p, wk.data(), r.data());
        for (i64 i = 0; i < p; i++) x[i] = -1.0 * r[i] + x[i];   // dgemv alpha=-1, beta=1
    }
    void step_z()
    {
        const double pen = 1.0 / s.rho;
        for (i64 i = 0; i < p; i++) {
            const double u = x[i] + s.adj_y[i] / s.rho;
            if (u > pen) s.z[i] = u - pen; else if (u < -pen) s.z[i] = u + pen; else s.z[i] = 0.0;
        }
    }
    double step_residual_and_dual()
    {
        for (i64 i = 0; i < p; i++) r[i] = x[i] - s.z[i];
        const double rp = norm2(r.data(), p);
        for (i64 i = 0; i < p; i++) s.y[i] = s.adj_y[i] + s.rho * r[i];
        return rp;
    }
    int solve(int maxit, Trace* tr) { return fast_admm_solve<double>(*this, s, maxit, true, tr); }
};

// lambda grid  (/root/reference/src/Lasso.cpp:78-89)
void make_lambda_grid(double lmax, double ratio, int nl, double* out)
{
    // Eigen's LinSpaced (3.3): a single point is `high`; else low + i*step with the last point forced to
    // high, or -- when |high| < |low| -- high - (nl-1-i)*step with the first point forced to low.
    const double a = std::log(lmax), b = std::log(ratio * lmax);
    if (nl == 1) { out[0] = std::exp(b); return; }
    const double h = (b - a) / (nl - 1);
    if (std::fabs(b) < std::fabs(a)) {
        out[0] = std::exp(a);
        for (int i = 1; i < nl; i++) out[i] = std::exp(b - (double)(nl - 1 - i) * h);
    } else {
        for (int i = 0; i < nl - 1; i++) out[i] = std::exp(a + (double)i * h);
        out[nl - 1] = std::exp(b);
    }
}

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// =====================================================================================
// C interface (ctypes).  All matrices column-major.  Return 0 on success.
// =====================================================================================
extern "C" {

int oracle_load_blas(const char* path) { return blas_load(path); }
int oracle_blas_threads(int nt)
{
    if (blas().on && blas().set_threads && nt > 0) blas().set_threads(nt);
    return (blas().on && blas().get_threads) ? blas().get_threads() : 0;
}
int oracle_omp_threads(int nt)
{
#ifdef _OPENMP
    if (nt > 0) omp_set_num_threads(nt);
    return omp_get_max_threads();
#else
    (void)nt; return 1;
#endif
}

// DataStd on float data, in place.  stats: meanY, scaleY; meanX/scaleX may be null.
int oracle_standardize_f32(float* X, float* Y, i64 n, i64 p, int standardize, int intercept,
                           float* meanX, float* scaleX, float* meanY_scaleY)
{
    Standardizer<float> st(n, p, standardize != 0, intercept != 0);
    st.apply(X, Y);
    if (meanX && !st.meanX.empty()) std::copy(st.meanX.begin(), st.meanX.end(), meanX);
    if (scaleX && !st.scaleX.empty()) std::copy(st.scaleX.begin(), st.scaleX.end(), scaleX);
    if (meanY_scaleY) { meanY_scaleY[0] = st.meanY; meanY_scaleY[1] = st.scaleY; }
    return 0;
}

// see g_lanczos_ncv; returns the previous value
int oracle_set_lanczos_ncv(int ncv) { const int old = g_lanczos_ncv; g_lanczos_ncv = ncv; return old; }

// coarse lambda_max of the symmetric matrix whose lower triangle is in S (n x n)
float oracle_coarse_eig_f32(const float* S, i64 n, int* info3)
{
    LanczosInfo li;
    float ev = coarse_largest_eigenvalue<float>([S, n](const float* v, float* w) { symv_lower(S, n, v, w); }, n, &li, 10, 0.1f, g_lanczos_ncv);
    if (info3) { info3[0] = li.nmatvec; info3[1] = li.nrestart; info3[2] = li.converged; }
    return ev;
}

void oracle_gram_tn_f32(const float* X, i64 n, i64 p, float* G) { gram_tn_lower(X, n, p, G); }

// Tall lasso/enet path starting from the Gram matrix of the *standardised* data.
//   G (p x p, lower valid, destroyed), XY = X'y (p), lambda_internal[nl] = lambda*n/scaleY.
//   z_out: nl x p floats (row k = solution on the standardised scale), niter_out[nl].
//   aux_out (optional, 4 doubles): rho, eigenvalue estimate, lanczos matvecs, setup seconds.
//   max_lambdas_timed: stop after this many lambdas if > 0 (bounded CPU-baseline sample).
int oracle_tall_path_from_gram(float* G, const float* XY, i64 p, int enet, double alpha,
                               const double* lambda_internal, int nl, int maxit,
                               double eps_abs, double eps_rel, double rho,
                               float* z_out, int* niter_out, double* aux_out,
                               double* trace, int trace_cap, int trace_lambda)
{
    TallLasso m(p, XY, eps_abs, eps_rel);
    if (enet) m.set_enet(alpha);
    std::vector<float> Gv(G, G + (size_t)p * p);
    double t0 = now_s();
    int info = m.init(Gv, lambda_internal[0], rho);
    double t1 = now_s();
    if (info != 0) return info < 0 ? info : 100 + info;
    for (int k = 0; k < nl; k++) {
        if (k > 0) m.init_warm(lambda_internal[k]);
        Trace tr; tr.buf = trace; tr.cap = trace_cap;
        niter_out[k] = m.solve(maxit, (trace && k == trace_lambda) ? &tr : nullptr);
        std::copy(m.s.z.begin(), m.s.z.end(), z_out + (size_t)k * p);
    }
    if (aux_out) { aux_out[0] = m.s.rho; aux_out[1] = m.ev_estimate; aux_out[2] = m.lz.nmatvec; aux_out[3] = t1 - t0; }
    return 0;
}

// The synthetic benchmark design (see synth_stream.hpp): rows [row0, row0 + nrows) of the global matrix
// into X (column-major, leading dimension nrows) and y.  Bit-identical to the library's CUDA generator.
int oracle_synth_f32(float* X, float* y, i64 nrows, i64 p, i64 row0, unsigned long long seed,
                     float mean_x, float sd_x, int nsig, float noise)
{
    synth::fill_x(X, nrows, nrows, p, row0, seed, mean_x, sd_x);
    if (y) synth::fill_y(X, nrows, nrows, p, row0, seed, nsig, noise, y);
    return 0;
}

// Full-size tall lasso on the synthetic design WITHOUT holding X (n x p floats = 40 GB at the headline
// size): the design is regenerated chunk by chunk for each of the three passes DataStd + Gram need
//   pass 1  column sums -> meanX; y as a whole -> meanY, scaleY            (DataStd.h:128-133, :98-104)
//   pass 2  centred sums of squares -> scaleX = |x - mean| / sqrt(n)       (DataStd.h:134-150)
//   pass 3  standardise the chunk, XY += Xc'y, G += Xc'Xc (SYRK)           (ADMMLassoTall.h:172,191-192)
// then the reference's lambda grid and warm-started path (Lasso.cpp:78-124) on (G, XY).
// Per-column statistics are float sums per chunk combined in double (the reference sums a whole column in
// float; the difference is summation order only).  Generation time is reported separately: X is the
// caller's input, not part of the reference's work.
//   times_out (8): generate, standardize (statistics + apply), gram (+ X'y), lanczos + cholesky, iterations, total, 0, 0
//   aux_out (6): rho, eig estimate, lambda0, scaleY, meanY, total iterations
int oracle_tall_fit_synth(i64 n, i64 p, unsigned long long seed, float mean_x, float sd_x, int nsig, float noise,
                          i64 chunk_rows, int enet, double alpha, double lambda_frac /* > 0: the single lambda = frac * lambda_max */,
                          int nlambda, double lmin_ratio, int maxit, double eps_abs, double eps_rel, double rho,
                          double* lambda_out, double* beta_out /* (p+1) x nl */, int* niter_out,
                          double* times_out, double* aux_out, float* gram_out /* p x p lower, optional */, float* xy_out /* optional */)
{
    if (!(n > p) || p < 3 || chunk_rows < 4) return -1;
    chunk_rows = (chunk_rows / 4) * 4;
    const double t_begin = now_s();
    double t_gen = 0, t_std = 0, t_gram = 0;
    std::vector<float> Xc((size_t)chunk_rows * p), yv(n);
    std::vector<double> colsum(p, 0.0), colsq(p, 0.0);
    std::vector<float> meanX(p), scaleX(p), invX(p);

    // ---- pass 1: y (needs the first nsig columns of every chunk) and column sums ----------------------------
    for (i64 r0 = 0; r0 < n; r0 += chunk_rows) {
        const i64 nr = std::min(chunk_rows, n - r0);
        double ta = now_s();
        synth::fill_x(Xc.data(), nr, nr, p, r0, seed, mean_x, sd_x);
        synth::fill_y(Xc.data(), nr, nr, p, r0, seed, nsig, noise, yv.data() + r0);
        double tb = now_s();
#pragma omp parallel for schedule(static)
        for (i64 j = 0; j < p; j++) colsum[j] += (double)(mean(Xc.data() + j * nr, nr) * (float)nr);
        t_gen += tb - ta; t_std += now_s() - tb;
    }
    double ta = now_s();
    for (i64 j = 0; j < p; j++) meanX[j] = (float)(colsum[j] / (double)n);
    const float n_invsqrt = float(1.0 / std::sqrt(float(n)));
    const float meanY = mean(yv.data(), n);
    for (i64 i = 0; i < n; i++) yv[i] -= meanY;
    const float scaleY = norm2(yv.data(), n) * n_invsqrt;
    for (i64 i = 0; i < n; i++) yv[i] /= scaleY;
    t_std += now_s() - ta;

    // ---- pass 2: centred sums of squares ----------------------------------------------------------------------
    for (i64 r0 = 0; r0 < n; r0 += chunk_rows) {
        const i64 nr = std::min(chunk_rows, n - r0);
        ta = now_s();
        synth::fill_x(Xc.data(), nr, nr, p, r0, seed, mean_x, sd_x);
        double tb = now_s();
#pragma omp parallel for schedule(static)
        for (i64 j = 0; j < p; j++) {
            const float* c = Xc.data() + j * nr;
            const float mu = meanX[j];
            float s0 = 0, s1 = 0, s2 = 0, s3 = 0; i64 i = 0;
            for (; i + 4 <= nr; i += 4) {
                const float a = c[i] - mu, b = c[i + 1] - mu, cc = c[i + 2] - mu, d = c[i + 3] - mu;
                s0 += a * a; s1 += b * b; s2 += cc * cc; s3 += d * d;
            }
            for (; i < nr; i++) { const float a = c[i] - mu; s0 += a * a; }
            colsq[j] += (double)((s0 + s1) + (s2 + s3));
        }
        t_gen += tb - ta; t_std += now_s() - tb;
    }
    for (i64 j = 0; j < p; j++) {
        scaleX[j] = (float)std::sqrt(colsq[j]) * n_invsqrt;
        invX[j] = float(1.0 / scaleX[j]);
    }

    // ---- pass 3: standardise, X'y, Gram ---------------------------------------------------------------------
    std::vector<float> G((size_t)p * p, 0.f), XY(p, 0.f), xyc(p);
    for (i64 r0 = 0; r0 < n; r0 += chunk_rows) {
        const i64 nr = std::min(chunk_rows, n - r0);
        ta = now_s();
        synth::fill_x(Xc.data(), nr, nr, p, r0, seed, mean_x, sd_x);
        double tb = now_s();
#pragma omp parallel for schedule(static)
        for (i64 j = 0; j < p; j++) {
            float* c = Xc.data() + j * nr;
            const float mu = meanX[j], inv = invX[j];
            for (i64 i = 0; i < nr; i++) c[i] = (c[i] - mu) * inv;
        }
        double tc = now_s();
        gemv_t(Xc.data(), nr, p, yv.data() + r0, xyc.data());
        for (i64 j = 0; j < p; j++) XY[j] += xyc[j];
        if (blas().on && fits_int(nr) && fits_int(p)) {
            int N = (int)p, K = (int)nr; float al = 1.f, be = 1.f;
            blas().ssyrk("L", "T", &N, &K, &al, Xc.data(), &K, &be, G.data(), &N);
        } else {
            std::vector<float> Gc((size_t)p * p, 0.f);
            gram_tn_lower_plain(Xc.data(), nr, p, Gc.data());
            for (size_t q = 0; q < G.size(); q++) G[q] += Gc[q];
        }
        double td = now_s();
        t_gen += tb - ta; t_std += tc - tb; t_gram += td - tc;
    }
    { std::vector<float>().swap(Xc); }
    if (gram_out) std::copy(G.begin(), G.end(), gram_out);
    if (xy_out) std::copy(XY.begin(), XY.end(), xy_out);

    // ---- lambda grid + path (Lasso.cpp:78-124) ------------------------------------------------------------
    TallLasso m(p, XY.data(), eps_abs, eps_rel);
    if (enet) m.set_enet(alpha);
    if (lambda_frac > 0) nlambda = 1;
    std::vector<double> lam(nlambda);
    const double lmax = (double)m.lambda0 / (double)n * (double)scaleY;
    if (lambda_frac > 0) lam[0] = lambda_frac * lmax;
    else make_lambda_grid(lmax, lmin_ratio, nlambda, lam.data());
    double t_setup = 0, t_iter = 0;
    long long total_it = 0;
    std::vector<float> coef(p);
    for (int k = 0; k < nlambda; k++) {
        const double il = lam[k] * (double)n / (double)scaleY;
        ta = now_s();
        if (k == 0) { const int info = m.init(G, il, rho); if (info != 0) return info < 0 ? info : 100 + info; }
        else m.init_warm(il);
        double tb = now_s();
        niter_out[k] = m.solve(maxit, nullptr);
        total_it += niter_out[k];
        t_setup += tb - ta; t_iter += now_s() - tb;
        coef = m.s.z;
        // DataStd::recover, flag 3 (DataStd.h:183-207): stored entries only
        float sacc = 0.f;
        for (i64 j = 0; j < p; j++) if (coef[j] != 0.f) { coef[j] /= scaleX[j]; coef[j] *= scaleY; sacc += coef[j] * meanX[j]; }
        double* col = beta_out + (size_t)k * (p + 1);
        col[0] = (double)(meanY - sacc);
        for (i64 j = 0; j < p; j++) col[j + 1] = coef[j];
        lambda_out[k] = lam[k];
    }
    if (times_out) {
        times_out[0] = t_gen; times_out[1] = t_std; times_out[2] = t_gram; times_out[3] = t_setup; times_out[4] = t_iter;
        times_out[5] = now_s() - t_begin; times_out[6] = 0; times_out[7] = 0;
    }
    if (aux_out) { aux_out[0] = m.s.rho; aux_out[1] = m.ev_estimate; aux_out[2] = m.lambda0; aux_out[3] = scaleY; aux_out[4] = meanY; aux_out[5] = (double)total_it; }
    return 0;
}

// Full entry: the reference's admm_lasso / admm_enet / admm_parlasso
// (/root/reference/src/Lasso.cpp:32-138, Enet.cpp:31-138, ParLasso.cpp:33-111).
//   x: n x p doubles (R layout), y: n doubles.  model: 0 lasso, 1 enet.  nthread > 1 -> consensus.
//   nlambda_in < 1 -> grid of `nlambda` values from lmin_ratio.
//   lambda_out[nl], beta_out[(p+1) x nl] column-major (row 0 = intercept), niter_out[nl].
//   aux_out (optional, 8 doubles): rho, eig estimate, lambda0, scaleY, meanY, t_std, t_setup, t_iter
int oracle_lasso_path(const double* x, const double* y, i64 n, i64 p, int model, double alpha,
                      const double* lambda_in, int nlambda_in, int nlambda, double lmin_ratio,
                      int standardize, int intercept, int nthread,
                      int maxit, double eps_abs, double eps_rel, double rho,
                      double* lambda_out, double* beta_out, int* niter_out, double* aux_out,
                      double* trace, int trace_cap, int trace_lambda)
{
    std::vector<float> X((size_t)n * p), Y(n);
    for (size_t i = 0; i < (size_t)n * p; i++) X[i] = (float)x[i];
    for (i64 i = 0; i < n; i++) Y[i] = (float)y[i];
    double t0 = now_s();
    Standardizer<float> st(n, p, standardize != 0, intercept != 0);
    st.apply(X.data(), Y.data());
    double t1 = now_s();

    const bool consensus = nthread > 1;
    const bool tall = n > p;
    std::vector<float> XY, G;
    TallLasso* mt = nullptr; WideLasso* mw = nullptr; ConsensusLasso* mc = nullptr;
    double lambda0 = 0;
    if (consensus) {
        mc = new ConsensusLasso(X.data(), Y.data(), n, p, nthread, eps_abs, eps_rel);
        lambda0 = mc->lambda0;
    } else if (tall) {
        XY.resize(p);
        gemv_t(X.data(), n, p, Y.data(), XY.data());
        mt = new TallLasso(p, XY.data(), eps_abs, eps_rel);
        if (model == 1) mt->set_enet(alpha);
        lambda0 = mt->lambda0;
    } else {
        mw = new WideLasso(X.data(), Y.data(), n, p, eps_abs, eps_rel);
        if (model == 1) mw->set_enet(alpha);
        lambda0 = mw->lambda0;
    }
    int nl = nlambda_in;
    std::vector<double> lam;
    if (nl < 1) {
        nl = nlambda;
        lam.resize(nl);
        const double lmax = lambda0 / n * double(st.scaleY);
        make_lambda_grid(lmax, lmin_ratio, nl, lam.data());
    } else lam.assign(lambda_in, lambda_in + nl);

    int rc = 0;
    double t_setup = 0, t_iter = 0;
    std::vector<float> coef(p);
    for (int k = 0; k < nl && rc == 0; k++) {
        const double il = lam[k] * n / double(st.scaleY);
        Trace tr; tr.buf = trace; tr.cap = trace_cap;
        Trace* trp = (trace && k == trace_lambda) ? &tr : nullptr;
        double ta = now_s();
        if (k == 0) {
            if (consensus) rc = mc->init(il, rho);
            else if (tall) {
                G.assign((size_t)p * p, 0.f);
                gram_tn_lower(X.data(), n, p, G.data());
                rc = mt->init(G, il, rho);
            } else mw->init(il, rho);
        } else {
            if (consensus) mc->init_warm(il); else if (tall) mt->init_warm(il); else mw->init_warm(il);
        }
        double tb = now_s();
        if (rc != 0) break;
        if (consensus) { niter_out[k] = mc->solve(maxit, trp); coef = mc->z; }
        else if (tall) { niter_out[k] = mt->solve(maxit, trp); coef = mt->s.z; }
        else { niter_out[k] = mw->solve(maxit, trp); coef = mw->x; }
        double tc = now_s();
        t_setup += tb - ta; t_iter += tc - tb;
        const float b0 = st.recover(coef.data());
        double* col = beta_out + (size_t)k * (p + 1);
        col[0] = b0;
        for (i64 j = 0; j < p; j++) col[j + 1] = coef[j];
        lambda_out[k] = lam[k];
    }
    if (aux_out) {
        aux_out[0] = consensus ? mc->rho : (tall ? mt->s.rho : mw->rho);
        aux_out[1] = consensus ? 0.0 : (tall ? mt->ev_estimate : (double)mw->sprad);
        aux_out[2] = lambda0; aux_out[3] = st.scaleY; aux_out[4] = st.meanY;
        aux_out[5] = t1 - t0; aux_out[6] = t_setup; aux_out[7] = t_iter;
    }
    delete mt; delete mw; delete mc;
    return rc;
}

// admm_lad (/root/reference/src/LAD.cpp:16-48).  beta_out: p+1 doubles (intercept first).
int oracle_lad(const double* x, const double* y, i64 n, i64 p, int intercept,
               int maxit, double eps_abs, double eps_rel, double rho,
               double* beta_out, int* niter_out, double* trace, int trace_cap)
{
    std::vector<double> X(x, x + (size_t)n * p), Y(y, y + n);
    Standardizer<double> st(n, p, true, intercept != 0);
    st.apply(X.data(), Y.data());
    LadModel m(X.data(), Y.data(), n, p, rho, eps_abs, eps_rel);
    Trace tr; tr.buf = trace; tr.cap = trace_cap;
    *niter_out = m.solve(maxit, trace ? &tr : nullptr);
    m.coefficients(beta_out + 1);
    // dense recover (DataStd.h:159-181): every entry is rescaled, zeros included
    std::vector<double> c(beta_out + 1, beta_out + 1 + p);
    double b0 = 0;
    if (st.flag == 1) { for (i64 j = 0; j < p; j++) { c[j] /= st.scaleX[j]; c[j] *= st.scaleY; } }
    else { double sacc = 0; for (i64 j = 0; j < p; j++) { c[j] /= st.scaleX[j]; c[j] *= st.scaleY; sacc += c[j] * st.meanX[j]; } b0 = st.meanY - sacc; }
    beta_out[0] = b0;
    std::copy(c.begin(), c.end(), beta_out + 1);
    return 0;
}

// admm_bp (/root/reference/src/BP.cpp:20-46).  beta_out: p doubles (zeros = not stored).
int oracle_bp(const double* x, const double* y, i64 n, i64 p,
              int maxit, double eps_abs, double eps_rel, double rho,
              double* beta_out, int* niter_out, double* trace, int trace_cap)
{
    BpModel m(x, y, n, p, rho, eps_abs, eps_rel);
    Trace tr; tr.buf = trace; tr.cap = trace_cap;
    *niter_out = m.solve(maxit, trace ? &tr : nullptr);
    std::copy(m.s.z.begin(), m.s.z.end(), beta_out);
    return 0;
}

}  // extern "C"
