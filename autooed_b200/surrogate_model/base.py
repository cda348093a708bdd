"""Surrogate-model base class with the surface and semantics of the reference's
`autooed/mobo/surrogate_model/base.py` (fit :31-52, evaluate :68-111, predict :142-149)."""
from abc import ABC, abstractmethod

import numpy as np

from autooed_b200.utils.normalization import StandardNormalization

DATA_TYPES = ('raw', 'continuous', 'normalized')  # how far along raw -> continuous -> normalised the caller's X already is


class SurrogateModel(ABC):
    def __init__(self, problem, **kwargs):
        self.problem = problem
        self.n_var, self.n_obj = problem.n_var, problem.n_obj
        self.bounds = np.array([problem.xl, problem.xu])
        self.transformation = problem.transformation
        self.normalization = StandardNormalization(self.bounds)
        self.fitted = False

    def _to_normalized(self, X, dtype, Y=None, refit=False):
        """Applies the remaining steps of raw -> continuous -> normalised to X (and Y when given); `refit` re-estimates
        the normalisation statistics first, which only `fit` does (base.py:44-50)."""
        if dtype == 'raw':
            X = self.transformation.do(X)
        if dtype != 'normalized':
            if refit:
                self.normalization.fit(X, Y)
            if Y is None:
                X = self.normalization.do(x=X)
            else:
                X, Y = self.normalization.do(x=X, y=Y)
        return X, Y

    def fit(self, X, Y, dtype='raw'):
        assert dtype in DATA_TYPES, f'Undefined data type {dtype} in surrogate fitting'
        Xn, Yn = self._to_normalized(X, dtype, Y, refit=True)
        self._fit(Xn, Yn)
        self.fitted = True

    @abstractmethod
    def _fit(self, X, Y):
        """X, Y normalised."""

    def evaluate(self, X, dtype='raw', std=False, gradient=False, hessian=False, calc_gradient=None,
                 calc_hessian=None):
        """Returns {'F','dF','hF','S','dS','hS'} (None when not requested), shapes as base.py:87-92.
        `calc_gradient` / `calc_hessian` (the DGEMO spelling) are accepted as aliases of `gradient` / `hessian`."""
        assert self.fitted, 'Surrogate model is not fitted yet'
        assert dtype in DATA_TYPES, f'Undefined data type {dtype} in surrogate evaluation'
        gradient = gradient if calc_gradient is None else calc_gradient
        hessian = hessian if calc_hessian is None else calc_hessian
        Xn, _ = self._to_normalized(X, dtype)
        return self._evaluate(Xn, std, gradient, hessian)

    @abstractmethod
    def _evaluate(self, X, std, gradient, hessian):
        """Evaluate on normalised inputs; implementations return UN-normalised outputs (the rescaling of
        base.py:106-109 is fused into the device epilogue)."""

    def predict(self, X, dtype='raw', std=False):
        out = self.evaluate(X, dtype=dtype, std=std)
        return (out['F'], out['S']) if std else out['F']
