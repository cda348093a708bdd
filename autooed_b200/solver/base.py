"""Solver base class — call surface of the reference's `autooed/mobo/solver/base.py:12-69`:
`solve(X, Y, batch_size, acquisition) -> (X_candidate, Y_candidate)`; the acquisition is wrapped into the solver-facing
SurrogateProblem and designs cross the raw <-> continuous boundary here."""
from abc import ABC, abstractmethod

from autooed_b200.surrogate_problem import SurrogateProblem


class Solver(ABC):
    def __init__(self, problem, **kwargs):
        self.real_problem, self.problem = problem, None  # `problem` becomes the surrogate problem of the current solve
        self.transformation = problem.transformation

    def solve(self, X, Y, batch_size, acquisition):
        self.problem = SurrogateProblem(self.real_problem, acquisition)
        Xc, Yc = self._solve(self.transformation.do(X), Y, batch_size)
        return self.transformation.undo(Xc), Yc

    @abstractmethod
    def _solve(self, X, Y, batch_size):
        """X continuous; returns candidate designs (continuous) and their acquisition values."""
