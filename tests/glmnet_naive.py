"""glmnet's Gaussian "naive" coordinate descent restated in NumPy (tests only).

The reference's README prints `range(coef(glmnet(x, y)) - admm$beta)` for its timing sections (README.md:225-241, :277-289).
glmnet is not a dependency of the reference's sources, only of that README chunk; its published algorithm (Friedman,
Hastie, Tibshirani 2010, "Regularization Paths for Generalized Linear Models via Coordinate Descent", J. Stat. Softw. 33(1),
and the strong rules of Tibshirani et al. 2012) is restated here the way glmnet 2.0's `elnet` runs it for nvars >= 500
(`type.gaussian = "naive"`), at glmnet's DEFAULT convergence threshold (thresh = 1e-7) -- so that the printed ranges, which
mix glmnet's own convergence error with the ADMM solvers', can be compared digit for digit:

  * observation weights 1/n; columns centred and scaled to unit (1/n) variance, y centred and scaled to unit variance;
  * lambda grid: lambda_1 = "infinity" (reported as lambda_2^2 / lambda_3), lambda_2 = flmin^(1/(nlam-1)) max_j |x_j'y| / alpha,
    then geometric with flmin = 0.01 (n < p) or 1e-4;
  * per lambda: strong-rule screening |x_j'r| > alpha (2 lambda - lambda_prev); cycles over the ever-active set until the
    largest weighted squared change of a cycle falls below thresh, then a cycle over the strong set, then a KKT check of the
    remaining variables (violators join the strong set and the cycle is repeated);
  * the very first lambda that needs work starts with a strong-set cycle, all later ones with the ever-active cycles;
  * path cut after at least five lambdas when the deviance ratio exceeds 0.999 or grows by less than 1e-5 of itself;
  * coefficients returned on the original scale with the intercept first.
"""
import numpy as np


def glmnet_gaussian_naive(x, y, alpha=1.0, nlambda=100, thresh=1e-7, maxit=100000):
    n, p = x.shape
    flmin = 0.01 if n < p else 1e-4
    w = 1.0 / n
    v = np.sqrt(w)
    xm = x.mean(axis=0)
    xc = (x - xm) * v
    xs = np.sqrt((xc * xc).sum(axis=0))
    X = np.asfortranarray(xc / xs)                     # unit columns: xv = 1
    ym = y.mean()
    r = (y - ym) * v
    ys = np.sqrt(r @ r)
    r = r / ys

    big, sml, rsqmax, mnlam, eps = 9.9e35, 1e-5, 0.999, 5, 1e-6
    bta, omb = alpha, 1.0 - alpha
    alf = max(eps, flmin) ** (1.0 / (nlambda - 1))
    a = np.zeros(p)
    inactive_ever = np.ones(p, dtype=bool)             # mm(k) == 0
    ia = []                                            # ever-active list in order of entry
    strong = np.zeros(p, dtype=bool)                   # ix
    g = np.abs(X.T @ r)
    rsq, nlp, iz, alm = 0.0, 0, 0, 0.0
    lam_out, coef_out, rsq_out = [], [], []

    def update(k, ab, dem):
        nonlocal rsq, r
        gk = X[:, k] @ r
        ak = a[k]
        u = gk + ak
        vv = abs(u) - ab
        new = np.copysign(vv, u) / (1.0 + dem) if vv > 0.0 else 0.0
        if new == ak:
            return 0.0
        a[k] = new
        if inactive_ever[k]:
            inactive_ever[k] = False
            ia.append(k)
        d = new - ak
        rsq += d * (2.0 * gk - d)
        r -= d * X[:, k]
        return d * d

    for m in range(nlambda):
        alm0 = alm
        if m == 0:
            alm = big
        elif m == 1:
            alm0 = g.max() / max(bta, 1e-3)
            alm = alf * alm0
        else:
            alm = alm * alf
        dem, ab, rsq0 = alm * omb, alm * bta, rsq
        jz = 1
        tlam = bta * (2.0 * alm - alm0)
        strong |= g > tlam
        while True:
            if not (iz * jz):
                nlp += 1
                dlx = 0.0
                for k in np.flatnonzero(strong):
                    dlx = max(dlx, update(k, ab, dem))
                if dlx < thresh:
                    rest = ~strong
                    g[rest] = np.abs(X[:, rest].T @ r)
                    viol = rest & (g > ab)
                    if viol.any():
                        strong |= viol
                        continue
                    break
                if nlp > maxit:
                    raise RuntimeError("glmnet: maxit reached")
            iz = 1
            while True:
                nlp += 1
                dlx = 0.0
                for k in list(ia):
                    dlx = max(dlx, update(k, ab, dem))
                if dlx < thresh:
                    break
                if nlp > maxit:
                    raise RuntimeError("glmnet: maxit reached")
            jz = 0
        lam_out.append(alm)
        coef_out.append(a.copy())
        rsq_out.append(rsq)
        if m + 1 < mnlam:
            continue
        if rsq - rsq0 < sml * rsq or rsq > rsqmax:
            break
    lam = np.array(lam_out) * ys
    lam[0] = np.exp(2.0 * np.log(lam[1]) - np.log(lam[2]))
    b = np.array(coef_out).T * ys / xs[:, None]
    a0 = ym - xm @ b
    return lam, np.vstack([a0[None, :], b]), np.array(rsq_out), nlp
