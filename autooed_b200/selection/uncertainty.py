"""Uncertainty selection — drop-in for `autooed/mobo/selection/uncertainty.py:14-19`: the batch_size candidates with
the largest product of predicted standard deviations (one fused device evaluation of the whole candidate set)."""
import numpy as np

from autooed_b200.selection.base import Selection


class Uncertainty(Selection):
    '''
    Selection based on uncertainty.
    '''
    def _select(self, X_candidate, Y_candidate, X, Y, batch_size):
        S = self.surrogate_model.evaluate(X_candidate, dtype='continuous', std=True)['S']
        order = np.argsort(np.prod(S, axis=1))[::-1]
        return X_candidate[order[:batch_size]]
