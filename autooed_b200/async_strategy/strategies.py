"""Asynchronous-batch strategies — SURVEY.md §8 f4, the surface of the reference's `autooed/mobo/async_strategy/`
(`base.py:8-16`, `kb.py:8-24`, `lp.py:9-23`, `bp.py:11-40`): how designs that are still being evaluated (`X_busy`) enter
the next proposal.  Each `fit(X, Y, X_busy)` returns the data and the acquisition the solver should then use.

Two upstream defects are repaired rather than reproduced, because the upstream classes cannot run with them:
`kb.py` uses `np` without importing numpy, and `bp.py:25-30` reads `Y_busy` before anything assigns it — here it is the
surrogate's posterior mean at the busy designs, which is what the Kriging-believer half of that strategy believes in.
"""
from abc import ABC, abstractmethod

import numpy as np


def _penalized(name, acquisition):
    from autooed_b200.factory import init_async_acquisition
    return init_async_acquisition(name, acquisition)


class AsyncStrategy(ABC):

    def __init__(self, surrogate_model, acquisition, **kwargs):
        self.surrogate_model, self.acquisition = surrogate_model, acquisition

    @abstractmethod
    def fit(self, X, Y, X_busy):
        pass


class KrigingBeliever(AsyncStrategy):
    """Treat the posterior mean at the busy designs as if it had been observed (kb.py:10-24)."""

    def fit(self, X, Y, X_busy):
        self.surrogate_model.fit(X, Y)
        X = np.vstack([X, X_busy])
        Y = np.vstack([Y, self.surrogate_model.predict(X_busy)])
        self.surrogate_model.fit(X, Y)
        self.acquisition.fit(X, Y)
        return X, Y, self.acquisition


class LocalPenalizer(AsyncStrategy):
    """Keep the data, penalise the acquisition around the busy designs (lp.py:11-23)."""

    def __init__(self, surrogate_model, acquisition, penalize_acq='lp', **kwargs):
        super().__init__(surrogate_model, acquisition)
        self.penalize_acq = penalize_acq

    def fit(self, X, Y, X_busy):
        self.surrogate_model.fit(X, Y)
        acquisition = _penalized(self.penalize_acq, self.acquisition)
        acquisition.fit(X, Y, X_busy)
        return X, Y, acquisition


class BelieverPenalizer(AsyncStrategy):
    """Believe the busy designs the model is sure about, penalise around the others (bp.py:11-40)."""

    def __init__(self, surrogate_model, acquisition, factor=0.2, penalize_acq='lp', **kwargs):
        super().__init__(surrogate_model, acquisition)
        self.factor, self.penalize_acq = factor, penalize_acq

    def fit(self, X, Y, X_busy):
        self.surrogate_model.fit(X, Y)
        Y_busy, Y_busy_std = self.surrogate_model.predict(X_busy, std=True)
        rel_std = self.surrogate_model.normalization.scale(y=Y_busy_std)  # in units of the data's spread (bp.py:23)
        believe_prob = np.maximum(1 - self.factor * rel_std, 0.0)
        believed = (np.random.uniform(size=Y_busy.shape) < believe_prob).all(axis=1)
        X, Y = np.vstack([X, X_busy[believed]]), np.vstack([Y, Y_busy[believed]])
        X_busy = X_busy[~believed] if (~believed).any() else None
        self.surrogate_model.fit(X, Y)
        acquisition = _penalized(self.penalize_acq, self.acquisition)
        acquisition.fit(X, Y, X_busy)
        return X, Y, acquisition
