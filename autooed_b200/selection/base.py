"""Selection base class — same surface as the reference's `autooed/mobo/selection/base.py:23-57`."""
from abc import ABC, abstractmethod

import numpy as np


class Selection(ABC):
    def __init__(self, surrogate_model, **kwargs):
        self.surrogate_model = surrogate_model
        self.transformation = surrogate_model.transformation

    def select(self, X_candidate, Y_candidate, X, Y, batch_size):
        # fill the candidates in case less than batch size (base.py:46-50)
        len_candidate = len(X_candidate)
        if len_candidate < batch_size:
            indices = np.concatenate([np.arange(len_candidate),
                                      np.random.choice(np.arange(len_candidate), batch_size - len_candidate)])
            X_candidate = X_candidate[indices]
            Y_candidate = Y_candidate[indices]
        X_candidate = self.transformation.do(X_candidate)
        X = self.transformation.do(X)
        X_next = self._select(X_candidate, Y_candidate, X, Y, batch_size)
        X_next = self.transformation.undo(X_next)
        return X_next

    @abstractmethod
    def _select(self, X_candidate, Y_candidate, X, Y, batch_size):
        pass
