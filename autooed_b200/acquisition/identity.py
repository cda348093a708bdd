"""Identity acquisition (reference: autooed/mobo/acquisition/identity.py:15-18): passes mean, gradient, hessian through."""
from autooed_b200 import _lib
from autooed_b200.acquisition.base import Acquisition


class Identity(Acquisition):
    kind = _lib.ACQ_IDENTITY

    def _fit(self, X, Y):
        pass
