"""Symmetric test matrices for the coarse-eigenvalue pins: Gram matrices from n >> p to n ~ p (the latter need up to
three implicit restarts at tol = 0.1), geometric and clustered spectra.  Deterministic."""
import numpy as np


def cases():
    rng = np.random.default_rng(0)
    out = []
    for n, p in ((2000, 50), (5000, 120), (300, 200), (120, 100), (64, 60), (400, 390), (90, 30), (1000, 500)):
        X = rng.normal(size=(n, p)).astype(np.float32)
        out.append(("gram %dx%d" % (n, p), X.T @ X))
    for n, p in ((50, 200), (100, 101)):                       # XX' of a wide matrix (the wide solver's gamma)
        X = rng.normal(size=(n, p)).astype(np.float32)
        out.append(("xxt %dx%d" % (n, p), X @ X.T))
    for p, decay in ((40, 0.9), (100, 0.97), (200, 0.99), (60, 0.5), (150, 0.995), (80, 0.8)):
        Q, _ = np.linalg.qr(rng.normal(size=(p, p)))
        d = decay ** np.arange(p)
        out.append(("geo p=%d %.3f" % (p, decay), ((Q * d) @ Q.T * 1000)))
    for p in (30, 80, 200):
        Q, _ = np.linalg.qr(rng.normal(size=(p, p)))
        d = np.ones(p)
        d[:5] = [2, 1.99, 1.98, 1.97, 1.96]
        out.append(("cluster p=%d" % p, (Q * d) @ Q.T))
    for p in (3, 4, 7):                                         # smallest sizes Spectra accepts (ncv = 3 <= n)
        X = rng.normal(size=(20, p)).astype(np.float32)
        out.append(("tiny p=%d" % p, X.T @ X))
    d = np.abs(rng.normal(size=64)) + 0.1
    out.append(("diag 64", np.diag(d)))
    return [(name, np.asfortranarray(((S + S.T) / 2).astype(np.float32))) for name, S in out]
