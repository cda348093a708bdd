"""The selection rules of the reference that need no search: `Uncertainty` (`autooed/mobo/selection/uncertainty.py`,
used by USeMO), `Direct` (`autooed/mobo/selection/direct.py`, used when the solver already returns one batch) and `Random`
(`autooed/mobo/selection/random.py`)."""
import numpy as np

from autooed_b200.selection.base import Selection


def most_uncertain(S, k):
    """Indices of the k rows of S (predicted standard deviations, one column per objective) with the largest product,
    in the reference's order: descending `argsort` of the products (uncertainty.py:17-18)."""
    volume = np.prod(S, axis=1)
    return np.argsort(volume)[::-1][:k]


class Uncertainty(Selection):
    """Keeps the batch_size candidates the surrogate is least sure about (one fused device evaluation of the set)."""

    def _select(self, X_candidate, Y_candidate, X, Y, batch_size):
        pred = self.surrogate_model.evaluate(X_candidate, dtype='continuous', std=True)
        return X_candidate[most_uncertain(pred['S'], batch_size)]


class Direct(Selection):
    """Passes the solver's candidates through; they must already be exactly one batch."""

    def _select(self, X_candidate, Y_candidate, X, Y, batch_size):
        n = len(X_candidate)
        assert n == batch_size, f'Candidate size {n} != batch size {batch_size}, cannot use direct selection'
        return X_candidate


class Random(Selection):
    """Uniformly random batch without replacement; consumes the global numpy RNG exactly as the reference (random.py:15-17)."""

    def _select(self, X_candidate, Y_candidate, X, Y, batch_size):
        return X_candidate[np.random.choice(len(X_candidate), size=batch_size, replace=False)]
