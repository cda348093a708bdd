"""Penalised acquisitions for asynchronous batches — SURVEY.md §8 f4, the surface of the reference's
`autooed/mobo/acquisition/penalize.py` (`PenalizedAcquisition` :11-75, `hammer_function` :83-87, Lipschitz estimates
:101-158, `LocalPenalization` :164-191, `LocalLipschitzPenalization` :194-202, `HardLocalPenalization` :205-227).

Designs that are still being evaluated (`X_busy`) carve exclusion zones out of the base acquisition.  Everything heavy
goes through the GPU surrogate: the 500-row gradient batches and the one-row L-BFGS probes of the Lipschitz estimate are
`GaussianProcess.evaluate(..., gradient=True)` calls, `predict(X_busy, std=True)` is the posterior kernel, the base
acquisition value is the fused acquisition call.  The zone arithmetic itself is (rows x n_busy x n_obj) host numpy.
These classes have no fused device transform (`kind` stays ACQ_NONE), so the NSGA-II solver drives them from its host loop.
"""
import numpy as np
from scipy.optimize import minimize
from scipy.stats import norm

from autooed_b200 import _lib
from autooed_b200.utils.blas import single_threaded_blas
from autooed_b200.utils.sampling import lhs

_DTYPES = ['raw', 'continuous', 'normalized']


def pairwise_distance(X, X0):
    """Euclidean distances (rows of X) x (rows of X0)  (penalize.py:78-80)."""
    diff = np.atleast_2d(X)[:, None, :] - np.atleast_2d(X0)[None, :, :]
    return np.sqrt((diff * diff).sum(axis=-1))


def _safe_ratio(num, den):
    """x / y with 0 where y == 0 (the reference's `safe_divide`)."""
    num, den = np.broadcast_arrays(np.asarray(num, dtype=float), np.asarray(den, dtype=float))
    out = np.zeros(num.shape)
    np.divide(num, den, out=out, where=den != 0)
    return out


def hammer_function(X, X0, R, S):
    """log of the product of the exclusion-zone factors Phi((|x - x_busy| - R) / S), per objective  (penalize.py:83-87)."""
    z = _safe_ratio(pairwise_distance(X, X0)[:, :, None] - R, S)
    return norm.logcdf(z).sum(axis=1)


def softplus(x):
    return np.log1p(np.exp(-np.abs(x))) + np.maximum(x, 0)


def neg_gradient_norm(X, surrogate_model):
    """-|d mu_i / dx| for every row and objective (penalize.py:90-92): one batched GPU gradient call."""
    out = surrogate_model.evaluate(X, dtype='continuous', std=False, gradient=True, hessian=False)
    return -np.linalg.norm(out['dF'], axis=-1)


def _neg_gradient_norm_of(x, surrogate_model, i):
    return neg_gradient_norm(np.atleast_2d(x), surrogate_model).ravel()[i]


def _steepest_point(surrogate_model, i, X_sample, bounds):
    """max over the box of |d mu_i / dx|: best sample, polished by L-BFGS-B (numerical gradient, like the reference)."""
    start = X_sample[np.argmin(neg_gradient_norm(X_sample, surrogate_model)[:, i])]
    with single_threaded_blas():
        res = minimize(_neg_gradient_norm_of, start, method='L-BFGS-B', bounds=bounds, args=(surrogate_model, i))
    L = -float(res.fun)
    return 10.0 if L < 1e-7 else L  # flat model: the reference's fallback constant


def estimate_lipschitz_constant(surrogate_model, n_sample=500):
    """Global Lipschitz constant of every objective's posterior mean over the unit cube (penalize.py:101-124)."""
    n_var, n_obj = surrogate_model.n_var, surrogate_model.n_obj
    X_sample = lhs(n_var, n_sample)
    bounds = np.column_stack([np.zeros(n_var), np.ones(n_var)])
    return np.array([_steepest_point(surrogate_model, i, X_sample, bounds) for i in range(n_obj)])


def estimate_local_lipschitz_constant(surrogate_model, X_busy, n_sample=500):
    """Lipschitz constants inside a half-length-scale box around each busy design, (n_busy, n_obj) (penalize.py:127-158)."""
    assert hasattr(surrogate_model, 'gps'), 'local Lipschitz constants need the GP length scales'
    n_var, n_obj = surrogate_model.n_var, surrogate_model.n_obj
    L = np.zeros((len(X_busy), n_obj))
    for i in range(n_obj):
        ell = np.asarray(surrogate_model.gps[i].kernel_.get_params()['k1__k2__length_scale'])
        lo, hi = np.maximum(X_busy - 0.5 * ell, 0.0), np.minimum(X_busy + 0.5 * ell, 1.0)
        for j in range(len(X_busy)):
            X_sample = lo[j] + lhs(n_var, n_sample) * (hi[j] - lo[j])
            L[j, i] = _steepest_point(surrogate_model, i, X_sample, np.column_stack([lo[j], hi[j]]))
    return L


class PenalizedAcquisition:
    """Wraps a base acquisition; without busy designs it is the base acquisition (penalize.py:11-75)."""
    kind = _lib.ACQ_NONE  # no fused device transform: solvers evaluate it through `evaluate`

    def __init__(self, base_acq):
        self.base_acq = base_acq
        self.surrogate_model = base_acq.surrogate_model
        self.transformation = base_acq.transformation
        self.X_busy = None

    @property
    def fitted(self):
        return self.base_acq.fitted

    def fit(self, X, Y, X_busy=None, dtype='raw'):
        assert dtype in _DTYPES, f'Undefined data type {dtype} in acquisition fitting'
        self.base_acq.fit(X, Y, dtype=dtype)
        if X_busy is None:
            return
        if dtype == 'raw':
            X, X_busy = self.transformation.do(X), self.transformation.do(X_busy)
        self.X_busy = np.asarray(X_busy, dtype=float)
        self._fit(X, Y)

    def evaluate(self, X, dtype='raw', gradient=False, hessian=False):
        assert dtype in _DTYPES, f'Undefined data type {dtype} in acquisition evaluation'
        if self.X_busy is None:
            return self.base_acq.evaluate(X, dtype, gradient, hessian)
        if dtype == 'raw':
            X = self.transformation.do(X)
        return self._evaluate(np.asarray(X, dtype=float), gradient, hessian)

    def _busy_posterior(self):
        return self.surrogate_model.predict(self.X_busy, dtype='continuous', std=True)

    def _base_value(self, X, gradient, hessian, who):
        if gradient or hessian:
            raise Exception(f'Gradient or hessian computation is not implemented yet in {who}.')
        return self.base_acq.evaluate(X, dtype='continuous', gradient=False, hessian=False)[0]


class LocalPenalization(PenalizedAcquisition):
    """Local penalisation (Gonzalez et al. 2016) with a global Lipschitz constant (penalize.py:164-191)."""

    def _lipschitz(self):
        return estimate_lipschitz_constant(self.surrogate_model)

    def _fit(self, X, Y):
        mean, std = self._busy_posterior()
        L = self._lipschitz()
        self.R = np.abs(mean - np.min(Y, axis=0)) / L  # zone radius: how far the believed value is from the best seen
        self.S = std / L

    def _evaluate(self, X, gradient, hessian):
        F = softplus(-self._base_value(X, gradient, hessian, 'Local Penalization'))
        # in log space the zones add up; back-transformed and negated because the solvers minimise
        return -np.exp(np.log(F) + hammer_function(X, self.X_busy, self.R, self.S)), None, None


class LocalLipschitzPenalization(LocalPenalization):
    """Local penalisation with one Lipschitz constant per busy design (penalize.py:194-202)."""

    def _lipschitz(self):
        return estimate_local_lipschitz_constant(self.surrogate_model, self.X_busy)


class HardLocalPenalization(PenalizedAcquisition):
    """Hard local penalisation (penalize.py:205-227; marked "TODO: check implementation" upstream, reproduced as is)."""

    def _fit(self, X, Y):
        mean, std = self._busy_posterior()
        L = estimate_lipschitz_constant(self.surrogate_model)
        self.R = (np.max(Y, axis=0) - mean) / L
        self.S = std / L

    def _evaluate(self, X, gradient, hessian):
        F = np.log1p(np.exp(-self._base_value(X, gradient, hessian, 'Hard Local Penalization')))
        with np.errstate(divide='ignore'):
            h = 1.0 / ((self.R + self.S)[None, :, :] * pairwise_distance(X, self.X_busy)[:, :, None]).prod(axis=1)  # (rows, n_obj)
        # soft minimum of (h, 1): the p = -5 norm of the pair, as the reference's `np.linalg.norm(..., -5, axis=0)`
        with np.errstate(divide='ignore'):
            clipped = (np.abs(h) ** -5.0 + 1.0) ** (-1.0 / 5.0)
        return F * -clipped, None, None
