"""Expected Improvement (reference: autooed/mobo/acquisition/ei.py:20-70)."""
import numpy as np

from autooed_b200 import _lib
from autooed_b200.acquisition.base import Acquisition


class ExpectedImprovement(Acquisition):
    kind = _lib.ACQ_EI

    def __init__(self, surrogate_model, **kwargs):
        super().__init__(surrogate_model, **kwargs)
        self.y_min = None

    def _fit(self, X, Y):
        self.y_min = np.min(Y, axis=0)  # ei.py:21

    def _param(self):
        return self.y_min
