"""Solver base class — same surface as the reference's `autooed/mobo/solver/base.py:12-69`: `solve(X, Y, batch_size,
acquisition)` wraps the acquisition into the solver-facing SurrogateProblem and converts raw <-> continuous."""
from abc import ABC, abstractmethod

from autooed_b200.surrogate_problem import SurrogateProblem


class Solver(ABC):
    def __init__(self, problem, **kwargs):
        self.real_problem = problem
        self.problem = None
        self.transformation = problem.transformation

    def solve(self, X, Y, batch_size, acquisition):
        self.problem = SurrogateProblem(self.real_problem, acquisition)
        X = self.transformation.do(X)
        X_candidate, Y_candidate = self._solve(X, Y, batch_size)
        X_candidate = self.transformation.undo(X_candidate)
        return X_candidate, Y_candidate

    @abstractmethod
    def _solve(self, X, Y, batch_size):
        pass
