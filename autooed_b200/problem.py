"""Minimal continuous problem description: the attributes the surrogate / acquisition / selection modules read from
an `autooed.problem.Problem` (`surrogate_model/base.py:24-28`, `surrogate_problem.py:27-30`).  A real AutoOED
`Problem` instance can be passed instead; this class only exists so the hot path is usable (and benchmarkable)
without the rest of AutoOED."""
import numpy as np


class ContinuousTransformation:
    """`autooed/problem/transformation.py:46-52`: raw <-> continuous is the identity for continuous variables."""

    def do(self, X):
        return np.array(X, dtype=object).astype(float)

    def undo(self, X):
        return np.array(X, dtype=float)


class SimpleProblem:
    def __init__(self, n_var, n_obj, xl=None, xu=None, n_constr=0, obj_type=None):
        self.n_var, self.n_obj, self.n_constr = int(n_var), int(n_obj), int(n_constr)
        self.xl = np.zeros(n_var) if xl is None else np.broadcast_to(np.asarray(xl, dtype=float), (n_var,)).copy()
        self.xu = np.ones(n_var) if xu is None else np.broadcast_to(np.asarray(xu, dtype=float), (n_var,)).copy()
        self.transformation = ContinuousTransformation()
        self.obj_type = ['min'] * n_obj if obj_type is None else obj_type

    def evaluate_constraint(self, x):
        return None
