"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's penalised acquisitions for asynchronous batches
(`autooed/mobo/acquisition/penalize.py`), on top of the GP oracle.  Pinned to the reference by
`tests/golden/async_*.npz` (oracle/make_golden_async.py).  Nothing in the product package may import this module.

`var_form` picks how the busy designs' posterior std is formed: 'kinv' is the reference's explicit-inverse arithmetic (what
the fixtures hold), 'chol' the triangular form the CUDA path uses.
"""
import numpy as np
from scipy.optimize import minimize
from scipy.special import log_ndtr

from oracle import gp_oracle as go
from oracle.ts_oracle import lhs  # the reference's sampler, same random-number consumption


def distance(X, X0):  # penalize.py:78-80
    X, X0 = np.atleast_2d(X), np.atleast_2d(X0)
    return np.sqrt(np.square(X[:, None, :] - X0[None, :, :]).sum(axis=-1))


def hammer(X, X0, R, S):  # penalize.py:83-87
    return log_ndtr(go._safe_div(distance(X, X0)[:, :, None] - R, S * np.ones_like(R))).sum(1)


def neg_grad_norm(orc, X):  # penalize.py:90-92
    return -np.linalg.norm(orc.evaluate(np.atleast_2d(X), std=False, gradient=True)["dF"], axis=-1)


def _polish(orc, i, X_sample, bounds):  # the body shared by penalize.py:114-122 and :149-156
    x0 = X_sample[np.argmin(neg_grad_norm(orc, X_sample)[:, i])]
    res = minimize(lambda x: neg_grad_norm(orc, x).flatten()[i], x0, method="L-BFGS-B", bounds=bounds)
    L = -float(res.fun)
    return 10.0 if L < 1e-7 else L


def lipschitz(orc, n_sample=500):  # penalize.py:101-124
    d = orc.n_var
    X_sample = lhs(d, n_sample)
    bounds = np.vstack([np.zeros(d), np.ones(d)]).T
    return np.array([_polish(orc, i, X_sample, bounds) for i in range(orc.n_obj)])


def local_lipschitz(orc, X_busy, n_sample=500):  # penalize.py:127-158
    d = orc.n_var
    L = np.zeros((len(X_busy), orc.n_obj))
    for i in range(orc.n_obj):
        ell = np.exp(orc.states[i].theta[1:-1])
        lo, hi = np.maximum(X_busy - 0.5 * ell, 0.0), np.minimum(X_busy + 0.5 * ell, 1.0)
        for j in range(len(X_busy)):
            X_sample = lo[j] + lhs(d, n_sample) * (hi[j] - lo[j])
            L[j, i] = _polish(orc, i, X_sample, np.vstack([lo[j], hi[j]]).T)
    return L


def fit_zones(orc, Y, X_busy, kind, var_form="kinv"):
    """(R, S) of LocalPenalization ('lp'), LocalLipschitzPenalization ('llp'), HardLocalPenalization ('hlp')."""
    out = orc.evaluate(X_busy, std=True, var_form=var_form)
    mean, std = out["F"], out["S"]
    L = local_lipschitz(orc, X_busy) if kind == "llp" else lipschitz(orc)
    R = (np.max(Y, axis=0) - mean) / L if kind == "hlp" else np.abs(mean - np.min(Y, axis=0)) / L
    return R, std / L


def softplus(x):
    return np.log1p(np.exp(-np.abs(x))) + np.maximum(x, 0)


def penalized_value(F_base, X, X_busy, R, S, kind):
    if kind in ("lp", "llp"):  # penalize.py:174-191
        return -np.exp(np.log(softplus(-F_base)) + hammer(X, X_busy, R, S))
    F = np.log1p(np.exp(-F_base))  # penalize.py:214-227
    with np.errstate(divide="ignore"):
        h = 1 / ((R + S)[None, :, :] * distance(X, X_busy)[:, :, None]).prod(axis=1).T
        clipped = np.vstack([np.linalg.norm(np.vstack([h[None, i], np.ones_like(h[None, i])]), -5, axis=0)
                             for i in range(h.shape[0])]).T
    return F * -clipped
