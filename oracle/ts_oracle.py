"""TEST INFRASTRUCTURE ONLY — CPU restatement of the Thompson-sampling acquisition of AutoOED
(`autooed/mobo/acquisition/ts.py`): random-Fourier-feature posterior sample (`_fit`, ts.py:26-63) and its value /
gradient / Hessian (`_evaluate`, ts.py:65-87), plus the LHS draw it consumes (`utils/sampling/lhs.py:126-144`).

Parity status: PINNED — `oracle/make_golden_ts.py` runs the unmodified reference class under a fixed numpy seed and
stores W, b, theta, sf2 and F / dF / hF in `tests/golden/ts_*.npz`; `tests/test_ts_oracle.py` replays them.
Only tests/, smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np
from scipy.stats import norm
from scipy.stats.distributions import chi2


def lhs(n, samples):
    """utils/sampling/lhs.py:126-144 (`_lhsclassic`): same random draws in the same order."""
    cut = np.linspace(0, 1, samples + 1)
    u = np.random.random((samples, n))
    pts = u * (cut[1:] - cut[:samples])[:, None] + cut[:samples, None]
    H = np.zeros_like(pts)
    for j in range(n):
        H[:, j] = pts[np.random.permutation(range(samples)), j]
    return H


def ts_fit(Xn, Yn, thetas, nu, M=100, mean_sample=False):
    """ts.py:26-63 after the per-objective GP refit (ts.py:36), whose result `thetas` is passed in
    (kernel log-parameters [log c1, log l, log c2] per objective).  Consumes the global numpy RNG like the reference.
    Returns lists (theta_feat, W, b, sf2) per objective."""
    n_sample, n_var = Xn.shape
    out = []
    for i, th in enumerate(thetas):
        ell = np.exp(th[1:-1])
        sf2 = np.exp(2 * th[0])   # ts.py:39 (note the factor 2: SURVEY.md B-7)
        sn2 = np.exp(2 * th[-1])  # ts.py:40
        sw1, sw2 = lhs(n_var, M), lhs(n_var, M)
        if nu > 0:
            W = np.tile(1. / ell, (M, 1)) * norm.ppf(sw1) * np.sqrt(nu / chi2.ppf(sw2, df=nu))  # ts.py:44
        else:
            W = np.random.uniform(size=(M, n_var)) * np.tile(1. / ell, (M, 1))  # ts.py:46
        b = 2 * np.pi * lhs(1, M)
        phi = np.sqrt(2. * sf2 / M) * np.cos(W @ Xn.T + np.tile(b, (1, n_sample)))
        A = phi @ phi.T + sn2 * np.eye(M)
        invcholA = np.linalg.inv(np.linalg.cholesky(A))
        invA = invcholA.T @ invcholA
        mu_theta = invA @ phi @ Yn[:, i]
        if mean_sample:
            theta = mu_theta
        else:
            cov = sn2 * invA
            cov = 0.5 * (cov + cov.T)
            theta = mu_theta + np.linalg.cholesky(cov) @ np.random.standard_normal(M)
        out.append((theta, W, b, sf2))
    return [o[0] for o in out], [o[1] for o in out], [o[2] for o in out], [o[3] for o in out]


def ts_evaluate(Xn, thetas, Ws, bs, sf2s, y_mean, y_scale, gradient=False, hessian=False):
    """ts.py:65-87 on normalised X: F un-normalised, dF rescaled by the Y scale, hF NOT rescaled (literal)."""
    Xn = np.atleast_2d(Xn)
    n = Xn.shape[0]
    F, dF, hF = [], [], []
    for theta, W, b, sf2 in zip(thetas, Ws, bs, sf2s):
        M = len(theta)
        factor = np.sqrt(2. * sf2 / M)
        t = W @ Xn.T + np.tile(b, (1, n))  # (M, n)
        F.append(factor * (theta @ np.cos(t)))
        if gradient:
            dF.append(-factor * (theta[None, :] * np.sin(t).T) @ W)
        if hessian:
            hF.append(-factor * np.einsum('ij,jk->ikj', theta[None, :] * np.cos(t).T, W) @ W)
    F = np.stack(F, axis=1) * y_scale + y_mean
    dF = np.stack(dF, axis=1) * np.asarray(y_scale)[None, :, None] if gradient else None
    hF = np.stack(hF, axis=1) if hessian else None
    return F, dF, hF
