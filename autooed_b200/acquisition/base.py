"""Acquisition base class — same surface as the reference's `autooed/mobo/acquisition/base.py`
(fit :26-45, evaluate :61-90).  Subclasses only declare which device-side transform to fuse."""
from abc import ABC, abstractmethod

from autooed_b200 import _lib

_DTYPES = ['raw', 'continuous', 'normalized']


class Acquisition(ABC):
    kind = _lib.ACQ_NONE

    def __init__(self, surrogate_model, exact=False, **kwargs):
        self.surrogate_model = surrogate_model
        self.transformation = surrogate_model.transformation
        self.normalization = surrogate_model.normalization
        self.exact = bool(exact)  # True: mathematically exact derivatives instead of the literal reference formulas
        self.fitted = False

    def fit(self, X, Y, dtype='raw'):
        assert self.surrogate_model.fitted, 'Surrogate model is not fitted before fitting acquisition function'
        assert dtype in _DTYPES, f'Undefined data type {dtype} in acquisition fitting'
        if dtype == 'raw':
            X = self.transformation.do(X)
        self._fit(X, Y)
        self.fitted = True

    @abstractmethod
    def _fit(self, X, Y):
        pass

    def _param(self):
        return None

    def evaluate(self, X, dtype='raw', gradient=False, hessian=False):
        """Returns (F, dF, hF): (N, n_obj), (N, n_obj, n_var), (N, n_obj, n_var, n_var); None when not requested."""
        assert self.fitted, 'Acquisition function is not fitted before evaluation'
        assert dtype in _DTYPES, f'Undefined data type {dtype} in acquisition evaluation'
        if dtype == 'raw':
            X = self.transformation.do(X)
            dtype = 'continuous'
        return self._evaluate(X, gradient, hessian, dtype=dtype)

    def _evaluate(self, X, gradient, hessian, dtype='continuous'):
        return self.surrogate_model.evaluate_acquisition(X, self.kind, self._param(), gradient=gradient,
                                                         hessian=hessian, exact=self.exact, dtype=dtype)
