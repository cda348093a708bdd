"""Upper Confidence Bound (reference: autooed/mobo/acquisition/ucb.py:19-53): F = mu - sqrt(log n / n) * sigma."""
from autooed_b200 import _lib
from autooed_b200.acquisition.base import Acquisition


class UpperConfidenceBound(Acquisition):
    kind = _lib.ACQ_UCB

    def __init__(self, surrogate_model, **kwargs):
        super().__init__(surrogate_model, **kwargs)
        self.n_sample = None

    def _fit(self, X, Y):
        self.n_sample = X.shape[0]  # ucb.py:20

    def _param(self):
        return [float(self.n_sample)]
