"""Independent NumPy restatement of the Spectra call the reference makes for its default rho / gamma --
TEST CODE ONLY, a third implementation next to the product's (admm_b200/csrc/coarse_eig.hpp) and the oracle's
(oracle/lanczos.hpp), which are textual siblings and therefore cannot pin each other.

    Spectra::SymEigsSolver<float, LARGEST_ALGE, DenseSymMatProd<float>> eigs(&op, 1, 3);
    eigs.init(); eigs.compute(10, 0.1);  ev = eigs.eigenvalues()[0]
    (/root/reference/src/ADMMLassoTall.h:196-201, src/ADMMLassoWide.h:202-207)

Written from the Spectra sources with matrices (as Eigen has them), not from either C++ restatement:
  SimpleRandom::random_vec            src/Spectra/SimpleRandom.h:38-76
  SymEigsSolver::init / compute       src/Spectra/SymEigsSolver.h:494-587
  factorize_from / restart            src/Spectra/SymEigsSolver.h:201-323
  num_converged / nev_adjusted        src/Spectra/SymEigsSolver.h:326-353
  retrieve_ritzpair                   src/Spectra/SymEigsSolver.h:356-397
  TridiagQR::compute / matrix_RQ / apply_YQ   src/Spectra/LinAlg/UpperHessenbergQR.h:337-371,467-598
The 3 x 3 tridiagonal eigenproblem (TridiagEigen.h) is handed to numpy.linalg.eigh in float64 and rounded to
float32 -- deliberately a different algorithm from the implicit-QR sweeps the C++ restatements carry, so that a
mis-stated sweep there shows up here.  Vector arithmetic is float32 throughout (Scalar = float).
"""
import numpy as np

F = np.float32


def simple_random_vec(n, seed):
    """SimpleRandom<float>(seed).random_vec(n): Lehmer generator a = 16807, m = 2^31 - 1 in the 16-bit split form."""
    a, m = 16807, 2147483647
    r = (seed & m) if seed else 1
    out = np.empty(n, dtype=F)
    for i in range(n):
        lo = a * (r & 0xFFFF)
        hi = a * (r >> 16)
        lo += (hi & 0x7FFF) << 16
        if lo > m:
            lo &= m
            lo += 1
        lo += hi >> 15
        if lo > m:
            lo &= m
            lo += 1
        r = lo
        out[i] = F(r) / F(m) - F(0.5)
    return out


def fnorm(v):
    return F(np.sqrt(np.dot(v, v), dtype=F))


class SymEigsLargest:
    def __init__(self, matvec, n, nev=1, ncv=3):
        if nev < 1 or nev > n - 1:
            raise ValueError("nev must satisfy 1 <= nev <= n - 1")
        if ncv <= nev or ncv > n:
            raise ValueError("ncv must satisfy nev < ncv <= n")
        self.op, self.n, self.nev, self.ncv = matvec, n, nev, ncv
        self.prec = F(np.power(np.finfo(F).eps, F(2.0) / F(3.0), dtype=F))
        self.nmatop = 0
        self.nrestart = 0

    def A(self, v):
        self.nmatop += 1
        return np.asarray(self.op(np.ascontiguousarray(v, dtype=F)), dtype=F)

    def init(self):
        n, m = self.n, self.ncv
        self.V = np.zeros((n, m), dtype=F)
        self.H = np.zeros((m, m), dtype=F)
        v = simple_random_vec(n, 0)
        nv = fnorm(v)
        if nv < self.prec:
            raise ValueError("initial residual vector cannot be zero")
        v = v / nv
        w = self.A(v)
        self.H[0, 0] = np.dot(v, w)
        self.f = w - v * self.H[0, 0]
        self.V[:, 0] = v

    def factorize_from(self, k, m, fk):
        if m <= k:
            return
        H, V = self.H, self.V
        self.f = fk.astype(F).copy()
        beta = fnorm(self.f)
        H[:, k:] = 0
        H[k:, :k] = 0
        for i in range(k, m):
            restart = False
            if beta < self.prec:
                self.f = simple_random_vec(self.n, 2 * i)
                Vi = V[:, :i]
                self.f = self.f - Vi @ (Vi.T @ self.f)
                beta = fnorm(self.f)
                restart = True
            v = self.f / beta
            V[:, i] = v
            H[i, i - 1] = F(0) if restart else beta
            w = self.A(v)
            Hii = F(np.dot(v, w))
            H[i - 1, i] = H[i, i - 1]
            H[i, i] = Hii
            if restart:
                self.f = w - Hii * v
            else:
                self.f = w - H[i, i - 1] * V[:, i - 1] - Hii * v
            beta = fnorm(self.f)
            Vi = V[:, :i + 1]
            Vf = Vi.T @ self.f
            count = 0
            while count < 5 and np.abs(Vf).max() > self.prec * beta:
                self.f = self.f - Vi @ Vf
                H[i - 1, i] += Vf[i - 1]
                H[i, i - 1] = H[i - 1, i]
                H[i, i] += Vf[i]
                beta = fnorm(self.f)
                Vf = Vi.T @ self.f
                count += 1

    def retrieve_ritzpair(self):
        T = np.diag(np.diag(self.H)).astype(np.float64)
        sub = np.diag(self.H, -1).astype(np.float64)
        T += np.diag(sub, -1) + np.diag(sub, 1)
        ev, U = np.linalg.eigh(T)
        order = np.argsort(-ev, kind="stable")                 # LARGEST_ALGE: descending
        self.ritz_val = ev[order].astype(F)
        self.ritz_est = U[self.ncv - 1, order].astype(F)        # last-row components of the Ritz vectors

    def num_converged(self, tol):
        thresh = F(tol) * np.maximum(np.abs(self.ritz_val[:self.nev]), self.prec)
        resid = np.abs(self.ritz_est[:self.nev]) * fnorm(self.f)
        self.ritz_conv = resid < thresh
        return int(self.ritz_conv.sum())

    def nev_adjusted(self, nconv):
        nev_new = self.nev + int((np.abs(self.ritz_est[self.nev:]) < self.prec).sum())
        nev_new += min(nconv, (self.ncv - nev_new) // 2)
        if nev_new == 1 and self.ncv >= 6:
            nev_new = self.ncv // 2
        elif nev_new == 1 and self.ncv > 2:
            nev_new = 2
        return nev_new

    @staticmethod
    def tridiag_qr(Hs):
        """Givens QR of the tridiagonal part of Hs: returns the rotations (c, s) and R (upper, two super-diagonals)."""
        m = Hs.shape[0]
        T = np.zeros((m, m), dtype=F)
        d, e = np.diag(Hs).astype(F), np.diag(Hs, -1).astype(F)
        T += np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        cs = []
        eps = np.finfo(F).eps
        for i in range(m - 1):
            a, b = T[i, i], T[i + 1, i]
            r = F(np.sqrt(a * a + b * b, dtype=F))
            if r <= eps:
                c, s, r = F(1), F(0), F(0)
            else:
                c, s = a / r, -b / r
            Gt = np.array([[c, -s], [s, c]], dtype=F)          # G' with G = [[c, s], [-s, c]]
            T[i:i + 2, i:] = Gt @ T[i:i + 2, i:]
            T[i, i], T[i + 1, i] = r, F(0)
            cs.append((c, s))
        return cs, T

    def restart(self, k):
        m = self.ncv
        if k >= m:
            return
        self.nrestart += 1
        Q = np.eye(m, dtype=F)
        for i in range(k, m):
            mu = self.ritz_val[i]
            Hs = self.H - mu * np.eye(m, dtype=F)
            cs, R = self.tridiag_qr(Hs)
            RQ = np.triu(R).copy()
            for j, (c, s) in enumerate(cs):                     # Y <- Y G_j on columns (j, j + 1), for Q and for R
                G = np.array([[c, s], [-s, c]], dtype=F)
                Q[:, j:j + 2] = Q[:, j:j + 2] @ G
                RQ[:, j:j + 2] = RQ[:, j:j + 2] @ G
            sub = np.diag(RQ, -1).copy()                        # matrix_RQ(): tridiagonal, super-diagonal := sub-diagonal
            Hn = np.diag(np.diag(RQ)) + np.diag(sub, -1) + np.diag(sub, 1)
            self.H = (Hn + mu * np.eye(m, dtype=F)).astype(F)
        Vs = (self.V @ Q[:, :k + 1]).astype(F)
        self.V[:, :k + 1] = Vs
        fk = self.f * Q[m - 1, k - 1] + self.V[:, k] * self.H[k, k - 1]
        self.factorize_from(k, m, fk)
        self.retrieve_ritzpair()

    def compute(self, maxit=10, tol=0.1):
        self.factorize_from(1, self.ncv, self.f)
        self.retrieve_ritzpair()
        nconv = 0
        for _ in range(maxit):
            nconv = self.num_converged(tol)
            if nconv >= self.nev:
                break
            self.restart(self.nev_adjusted(nconv))
        self.nconv = nconv
        return min(self.nev, nconv)


def coarse_largest_eigenvalue(S):
    """ev = eigs.eigenvalues()[0] for the symmetric float32 matrix S, with the matvec count and restart count."""
    S = np.asarray(S, dtype=F)
    e = SymEigsLargest(lambda v: S @ v, S.shape[0], 1, 3)
    e.init()
    e.compute(10, 0.1)
    return float(e.ritz_val[0]), e.nmatop, e.nrestart, int(e.nconv >= 1)
