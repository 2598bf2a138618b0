"""Regenerate the input data of the reference README's worked examples without R.

The reference has no tests and no fixtures; the only results it ever published are the
coefficient columns printed in /root/reference/README.md (lines 66-88 lasso + parallel lasso,
101-122 elastic net, 140-160 LAD, 180-182 basis pursuit).  Those columns are committed in
``readme_vectors.py``; this script rebuilds the *inputs* (``set.seed(123)`` + ``runif`` /
``rnorm`` / ``sample``) by restating R's documented generators:

* ``set.seed``: Mersenne-Twister seeded through the 69069 LCG scrambler (R: src/main/RNG.c);
* ``unif_rand``: MT19937 ``genrand_int32 * 2.3283064365386963e-10`` with the (0,1) fix-up;
* ``norm_rand`` (INVERSION): ``u = floor(2^27 u1) + u2; qnorm(u / 2^27)``;
* ``sample`` (R < 3.6.0, the README dates from 2015): partial Fisher-Yates with
  ``j = floor(m * unif_rand())``.

Known-answer check built in: seed 123 must give runif -> 0.2875775201246142 ... and
rnorm -> -0.56047564655221 ... (the values every R user has seen).

Run:  python tests/golden/make_readme_data.py      (writes the .npz files next to this script)
"""
import os
import numpy as np
from scipy.special import ndtri

HERE = os.path.dirname(os.path.abspath(__file__))


class RRng:
    def __init__(self, seed: int):
        s = np.uint32(seed)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(69069) * s + np.uint32(1)
            st = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = np.uint32(69069) * s + np.uint32(1)
                st[j] = s
        self.bg = np.random.MT19937()
        state = self.bg.state
        state["state"]["key"] = st[1:].copy()
        state["state"]["pos"] = 624
        self.bg.state = state

    def unif(self) -> float:
        v = float(self.bg.random_raw()) * 2.3283064365386963e-10
        i2_32m1 = 2.328306437080797e-10
        if v <= 0.0:
            return 0.5 * i2_32m1
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * i2_32m1
        return v

    def runif(self, k):
        return np.array([self.unif() for _ in range(k)])

    def norm(self) -> float:
        big = 134217728.0
        u = self.unif()
        u = float(int(big * u)) + self.unif()
        return float(ndtri(u / big))

    def rnorm(self, k, mean=0.0, sd=1.0):
        return np.array([mean + sd * self.norm() for _ in range(k)])

    # vectorised forms of unif / norm (the README's benchmark sections draw up to 1e7 normals)
    def runif_vec(self, k):
        v = self.bg.random_raw(k).astype(np.float64) * 2.3283064365386963e-10
        i2_32m1 = 2.328306437080797e-10
        v[v <= 0.0] = 0.5 * i2_32m1
        v[1.0 - v <= 0.0] = 1.0 - 0.5 * i2_32m1
        return v

    def rnorm_vec(self, k, mean=0.0, sd=1.0):
        u = self.runif_vec(2 * k)
        big = 134217728.0
        return mean + sd * ndtri((np.floor(big * u[0::2]) + u[1::2]) / big)

    def sample_old(self, v):
        v = np.asarray(v)
        m = len(v)
        idx = list(range(m))
        out = np.empty(m, dtype=v.dtype)
        for i in range(len(v)):
            j = int(m * self.unif())
            out[i] = v[idx[j]]
            m -= 1
            idx[j] = idx[m]
        return out


def lasso_data():
    r = RRng(123)
    n, p, m = 100, 20, 5
    b = np.concatenate([r.runif(m), np.zeros(p - m)])
    x = r.rnorm(n * p, 1.2, 2.0).reshape(p, n).T.copy(order="F")  # column-major fill
    y = 5.0 + x @ b + r.rnorm(n)
    return x, y, b


def bp_data():
    r = RRng(123)
    n, p, nsig = 50, 100, 15
    beta_true = np.concatenate([r.runif(nsig), np.zeros(p - nsig)])
    beta_true = r.sample_old(beta_true)
    x = r.rnorm(n * p).reshape(p, n).T.copy(order="F")
    y = x @ beta_true
    return x, y, beta_true


def benchmark_data(n, p, m=100):
    """Input of the README's timing sections (README.md:193-201 n > p, :246-254 p > n): set.seed(123);
    b <- c(runif(m), 0 ...); x <- matrix(rnorm(n * p, sd = 2), n, p); y <- x %*% b + rnorm(n).  Generated on the fly
    (80 MB at n = 1e4, p = 1e3), never stored."""
    r = RRng(123)
    b = np.concatenate([r.runif_vec(m), np.zeros(p - m)])
    x = r.rnorm_vec(n * p, 0.0, 2.0).reshape(p, n).T.copy(order="F")
    y = x @ b + r.rnorm_vec(n)
    return x, y, b


def bp_benchmark_data(n, p, nsig):
    """README.md:365-373 / :395-403: beta_true <- sample(c(runif(nsig), 0 ...)); x <- matrix(rnorm(n * p), n, p); y <- x %*% beta_true."""
    r = RRng(123)
    bt = np.concatenate([r.runif_vec(nsig), np.zeros(p - nsig)])
    bt = r.sample_old(bt)
    x = r.rnorm_vec(n * p).reshape(p, n).T.copy(order="F")
    return x, x @ bt, bt


def lad_benchmark_data(n, p):
    """README.md:299-305: b <- runif(p); x <- matrix(rnorm(n * p, sd = 2), n, p); y <- x %*% b + rnorm(n)."""
    r = RRng(123)
    b = r.runif_vec(p)
    x = r.rnorm_vec(n * p, 0.0, 2.0).reshape(p, n).T.copy(order="F")
    y = x @ b + r.rnorm_vec(n)
    return x, y, b


def main():
    r = RRng(123)
    assert np.allclose(r.runif_vec(3), [0.2875775201246142, 0.7883051354438066, 0.4089769218116999], atol=1e-15)
    r = RRng(123)
    assert np.allclose(r.rnorm_vec(3), [-0.56047564655221, -0.23017748948328, 1.55870831414912], atol=1e-12)
    r = RRng(123)
    u = r.runif(3)
    assert np.allclose(u, [0.2875775201246142, 0.7883051354438066, 0.4089769218116999], atol=1e-15), u
    r = RRng(123)
    z = r.rnorm(3)
    assert np.allclose(z, [-0.56047564655221, -0.23017748948328, 1.55870831414912], atol=1e-12), z
    x, y, b = lasso_data()
    np.savez(os.path.join(HERE, "readme_lasso_data.npz"), x=x, y=y, b=b)
    x, y, bt = bp_data()
    np.savez(os.path.join(HERE, "readme_bp_data.npz"), x=x, y=y, beta_true=bt)
    print("wrote readme_lasso_data.npz, readme_bp_data.npz")


if __name__ == "__main__":
    main()
