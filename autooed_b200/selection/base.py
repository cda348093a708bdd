"""Selection base class — same call surface as the reference's `autooed/mobo/selection/base.py:23-57`:
`select(X_candidate, Y_candidate, X, Y, batch_size) -> X_next` in raw design space."""
from abc import ABC, abstractmethod

import numpy as np


def pad_candidates(X_candidate, Y_candidate, batch_size):
    """Fewer candidates than the batch: repeat randomly drawn ones until there are batch_size rows (base.py:46-50; the
    draw uses the global numpy RNG exactly like the reference, so a fixed seed gives the same padding)."""
    n = len(X_candidate)
    if n >= batch_size:
        return X_candidate, Y_candidate
    extra = np.random.choice(np.arange(n), batch_size - n)
    rows = np.concatenate([np.arange(n), extra])
    return X_candidate[rows], Y_candidate[rows]


class Selection(ABC):
    def __init__(self, surrogate_model, **kwargs):
        self.surrogate_model = surrogate_model
        self.transformation = surrogate_model.transformation

    def select(self, X_candidate, Y_candidate, X, Y, batch_size):
        X_candidate, Y_candidate = pad_candidates(X_candidate, Y_candidate, batch_size)
        to_cont, to_raw = self.transformation.do, self.transformation.undo
        chosen = self._select(to_cont(X_candidate), Y_candidate, to_cont(X), Y, batch_size)  # continuous space
        return to_raw(chosen)

    @abstractmethod
    def _select(self, X_candidate, Y_candidate, X, Y, batch_size):
        """Returns the selected rows of X_candidate (continuous representation)."""
