"""Thompson sampling — drop-in for `autooed/mobo/acquisition/ts.py` (the acquisition of TSEMO, the reference's default
algorithm): a posterior function sample per objective through random Fourier features.

`_fit` keeps the reference's host arithmetic and, deliberately, its exact consumption of the global numpy RNG
(ts.py:26-63: two LHS designs for the spectral points, one for the phases, one normal draw for the weights, per
objective), so a seeded run draws the same sample; the M x M solves are tiny (M = 100).  `_evaluate` — the part the
solver calls for whole populations — runs on the GPU (`aoed_rff_eval`)."""
import numpy as np
from scipy.stats import norm
from scipy.stats.distributions import chi2

from autooed_b200 import _lib
from autooed_b200.acquisition.base import Acquisition
from autooed_b200.utils.sampling import lhs
from autooed_b200.utils.blas import single_threaded_blas


class ThompsonSampling(Acquisition):
    '''
    Thompson Sampling.
    '''
    def __init__(self, surrogate_model, n_spectral_pts=100, mean_sample=False, **kwargs):
        super().__init__(surrogate_model)
        assert hasattr(surrogate_model, 'gps') and hasattr(surrogate_model, 'nu'), \
            'Thompson Sampling requires Gaussian Process as the surroagte model'
        self.M = int(n_spectral_pts)
        self.mean_sample = mean_sample
        self.thetas, self.Ws, self.bs, self.sf2s = None, None, None, None

    def _fit(self, X, Y):
        Xn, Yn = self.normalization.do(x=X, y=Y)
        sm = self.surrogate_model
        n_sample, n_var, nu, M = Xn.shape[0], sm.n_var, sm.nu, self.M
        self.thetas, self.Ws, self.bs, self.sf2s = [], [], [], []
        with single_threaded_blas():
            self._fit_objectives(sm, Xn, Yn, n_var, nu, M)

    def _fit_objectives(self, sm, Xn, Yn, n_var, nu, M):
        for i, gp in enumerate(sm.gps):
            gp.fit(Xn, Yn[:, i])  # ts.py:36: per-objective refit on the normalised data
            log_par = gp.kernel_.theta
            inv_ell = 1. / np.exp(log_par[1:-1])  # 1/ell as the reference forms it (bit-identical spectral points)
            sf2, sn2 = np.exp(2 * log_par[0]), np.exp(2 * log_par[-1])  # ts.py:39-40 (squares of c1, c2: SURVEY.md B-7)
            u1, u2 = lhs(n_var, M), lhs(n_var, M)
            if nu > 0:  # Student-t spectral density of the Matern kernel (ts.py:44)
                W = inv_ell[None, :] * norm.ppf(u1) * np.sqrt(nu / chi2.ppf(u2, df=nu))
            else:       # ts.py:46
                W = np.random.uniform(size=(M, n_var)) * inv_ell[None, :]
            b = 2 * np.pi * lhs(1, M)
            phi = np.sqrt(2. * sf2 / M) * np.cos(W @ Xn.T + b)  # (M, n_sample)
            A = phi @ phi.T + sn2 * np.eye(M)
            half = np.linalg.inv(np.linalg.cholesky(A))
            invA = half.T @ half
            weights = invA @ phi @ Yn[:, i]
            if not self.mean_sample:
                cov = sn2 * invA
                weights = weights + np.linalg.cholesky(0.5 * (cov + cov.T)) @ np.random.standard_normal(M)
            self.thetas.append(weights)
            self.Ws.append(W)
            self.bs.append(b)
            self.sf2s.append(sf2)

    def device_sample(self):
        """The fitted sample as the contiguous arrays the C ABI takes: (W, b, theta, factor, y_mean, y_scale)."""
        m, M = len(self.thetas), self.M
        ys = self.normalization.y_scaler
        return (_lib.as_f64(np.stack(self.Ws)), _lib.as_f64(np.stack(self.bs).reshape(m, M)), _lib.as_f64(np.stack(self.thetas)),
                _lib.as_f64(np.sqrt(2. * np.asarray(self.sf2s) / M)), _lib.as_f64(ys.mean_), _lib.as_f64(ys.scale_))

    def _evaluate(self, X, gradient, hessian, dtype='continuous'):
        if dtype != 'normalized':
            X = self.normalization.do(x=X)
        X = _lib.as_f64(X)
        if X.ndim != 2 or X.shape[1] != self.surrogate_model.n_var:
            raise ValueError(f'X must be 2-D with {self.surrogate_model.n_var} columns, got shape {X.shape}')
        lib = _lib.load()
        _lib.require_device()
        C, d, m, M = X.shape[0], X.shape[1], len(self.thetas), self.M
        W, b, theta, factor, y_mean, y_scale = self.device_sample()
        F = _lib.empty((C, m))
        dF = _lib.empty((C, m, d)) if gradient else None
        hF = _lib.empty((C, m, d, d)) if hessian else None
        flags = (_lib.EVAL_GRAD if gradient else 0) | (_lib.EVAL_HESS if hessian else 0)
        device = getattr(self.surrogate_model, 'device', 0)
        _lib.check(lib.aoed_rff_eval(device, C, d, m, M, _lib.ptr(W), _lib.ptr(b), _lib.ptr(theta), _lib.ptr(factor),
                                     _lib.ptr(y_mean), _lib.ptr(y_scale), _lib.ptr(X), flags,
                                     _lib.ptr(F), _lib.ptr(dF), _lib.ptr(hF)))
        return F, dF, hF
