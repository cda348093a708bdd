"""The reverse-communication L-BFGS-B driver (autooed_b200/surrogate_model/lbfgsb_lockstep.py) against
`scipy.optimize.minimize(method='L-BFGS-B', jac=True)` — the call sklearn makes for the reference's GP fit
(_gpr.py `_constrained_optimization`): identical iterates, so identical x, f and evaluation counts, bit for bit."""
import numpy as np
import pytest
from scipy.optimize import minimize

from autooed_b200.surrogate_model import lbfgsb_lockstep as ll


def _rosen_like(seed, n):
    rng = np.random.default_rng(seed)
    a = rng.uniform(0.5, 2.0, n)

    def fg(x):
        f = np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (a[:-1] - x[:-1]) ** 2) + np.sum(np.cos(3 * x))
        g = np.zeros(n)
        g[:-1] += -400.0 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2.0 * (a[:-1] - x[:-1])
        g[1:] += 200.0 * (x[1:] - x[:-1] ** 2)
        return f, g - 3 * np.sin(3 * x)
    return fg


def test_self_check_passes_with_this_scipy():
    assert ll.available()


@pytest.mark.parametrize("n", [2, 8, 22])
def test_same_iterates_as_scipy_minimize(n):
    rng = np.random.default_rng(n)
    fgs = [_rosen_like(s, n) for s in range(7)]
    bounds = np.stack([np.full(n, -1.5), np.full(n, 2.0)], axis=1)
    bounds[0] = [0.3, 0.6]  # an active bound
    x0s = [rng.uniform(-2.0, 2.5, n) for _ in fgs]  # some starts outside the box: clipped like scipy does
    calls = []

    def batch(ids, X):
        calls.append(len(ids))
        out = [fgs[b](x) for b, x in zip(ids, X)]
        return np.array([o[0] for o in out]), np.stack([o[1] for o in out])

    got, n_batches, n_evals = ll.minimize_lockstep(batch, x0s, bounds)
    assert n_batches == len(calls) and n_evals == sum(calls) and calls[0] == len(fgs)
    for fg, x0, (x, f, nfev) in zip(fgs, x0s, got):
        ref = minimize(fg, x0, method="L-BFGS-B", jac=True, bounds=bounds)
        assert np.array_equal(ref.x, x) and ref.fun == f and ref.nfev == nfev
    assert n_batches == max(r[2] for r in got)  # lock-step: as many rounds as the longest run needs
