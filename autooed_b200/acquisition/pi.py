"""Probability of Improvement (reference: autooed/mobo/acquisition/pi.py:20-64)."""
import numpy as np

from autooed_b200 import _lib
from autooed_b200.acquisition.base import Acquisition


class ProbabilityOfImprovement(Acquisition):
    kind = _lib.ACQ_PI

    def __init__(self, surrogate_model, **kwargs):
        super().__init__(surrogate_model, **kwargs)
        self.y_min = None

    def _fit(self, X, Y):
        self.y_min = np.min(Y, axis=0)  # pi.py:21

    def _param(self):
        return self.y_min
