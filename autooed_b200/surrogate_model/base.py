"""Surrogate-model base class: same surface and semantics as the reference's
`autooed/mobo/surrogate_model/base.py` (fit :31-52, evaluate :68-111, predict :142-149)."""
from abc import ABC, abstractmethod

import numpy as np

from autooed_b200.utils.normalization import StandardNormalization

_DTYPES = ['raw', 'continuous', 'normalized']


class SurrogateModel(ABC):
    def __init__(self, problem, **kwargs):
        self.problem = problem
        self.n_var, self.n_obj = problem.n_var, problem.n_obj
        self.bounds = np.array([problem.xl, problem.xu])
        self.transformation = problem.transformation
        self.normalization = StandardNormalization(self.bounds)
        self.fitted = False

    def fit(self, X, Y, dtype='raw'):
        assert dtype in _DTYPES, f'Undefined data type {dtype} in surrogate fitting'
        if dtype == 'raw':
            X = self.transformation.do(X)
        if dtype == 'raw' or dtype == 'continuous':
            self.normalization.fit(X, Y)
            X, Y = self.normalization.do(x=X, y=Y)
        self._fit(X, Y)
        self.fitted = True

    @abstractmethod
    def _fit(self, X, Y):
        pass

    def evaluate(self, X, dtype='raw', std=False, gradient=False, hessian=False, calc_gradient=None,
                 calc_hessian=None):
        """Returns {'F','dF','hF','S','dS','hS'} (None when not requested), shapes as base.py:87-92.
        `calc_gradient` / `calc_hessian` are accepted as aliases of `gradient` / `hessian`."""
        assert self.fitted, 'Surrogate model is not fitted yet'
        assert dtype in _DTYPES, f'Undefined data type {dtype} in surrogate evaluation'
        if calc_gradient is not None:
            gradient = calc_gradient
        if calc_hessian is not None:
            hessian = calc_hessian
        if dtype == 'raw':
            X = self.transformation.do(X)
        if dtype == 'raw' or dtype == 'continuous':
            X = self.normalization.do(x=X)
        return self._evaluate(X, std, gradient, hessian)

    @abstractmethod
    def _evaluate(self, X, std, gradient, hessian):
        """Evaluate on normalised inputs; implementations return UN-normalised outputs (the rescaling of
        base.py:106-109 is fused into the device epilogue)."""
        pass

    def predict(self, X, dtype='raw', std=False):
        out = self.evaluate(X, dtype=dtype, std=std)
        if std:
            return out['F'], out['S']
        return out['F']
