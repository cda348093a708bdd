"""Solver-facing adapter with the call surface of the reference's `autooed/mobo/surrogate_problem.py:11-193`
(`SurrogateProblem.evaluate / evaluate_constraint`), i.e. the pymoo 0.4.1 `Problem` protocol that NSGA-II / MOEA-D /
ParetoDiscovery drive.  The acquisition it wraps evaluates on the GPU; this class is host logic only (which derivatives a
`return_values_of` list implies, the cheap-constraint loop, constraint violation) and does not need pymoo.
INTEGRATION.md shows how a maintainer re-bases it on `pymoo.model.problem.Problem`."""
import numpy as np


def calc_constraint_violation(G):
    """pymoo 0.4.1 `Problem.calc_constraint_violation`: row sums of the positive parts, as a column."""
    if G is None:
        return None
    if G.shape[1] == 0:
        return np.zeros((G.shape[0], 1))
    return np.maximum(G, 0.0).sum(axis=1, keepdims=True)


class SurrogateProblem:

    def __init__(self, problem, acquisition):
        self.problem, self.acquisition = problem, acquisition
        self.transformation = problem.transformation
        self.n_var, self.n_obj, self.n_constr = problem.n_var, problem.n_obj, getattr(problem, 'n_constr', 0)
        self.xl, self.xu = np.asarray(problem.xl, dtype=float), np.asarray(problem.xu, dtype=float)
        self.callback = None
        # pymoo's evaluation_of="auto": what `_evaluate` provides without being asked for derivatives
        self.evaluation_of = ['F', 'G'] if self.n_constr > 0 else ['F']

    def _evaluate(self, X, out, *args, gradient, hessian, **kwargs):
        """Acquisition on the continuous designs + the real problem's cheap constraints row by row
        (surrogate_problem.py:32-52)."""
        F, dF, hF = self.acquisition.evaluate(X, dtype='continuous', gradient=gradient, hessian=hessian)
        out.update(F=F, dF=dF, hF=hF)
        if self.n_constr == 0:  # CV is zero by definition (:136-137): skip the per-row Python loop (10^6 rows in the c4 case)
            out['G'] = None
        else:
            out['G'] = np.array([self.problem.evaluate_constraint(x) for x in self.transformation.undo(X)])

    def evaluate(self, X, *args, return_values_of="auto", return_as_dictionary=False, **kwargs):
        """surrogate_problem.py:54-166: accepts one design (1-D) or a batch (2-D)."""
        assert self.acquisition.fitted, 'Acquisition function is not fitted before surrogate problem evaluation'
        if self.callback is not None:
            self.callback(X)
        single = np.ndim(X) == 1
        X = np.atleast_2d(X)
        if X.shape[1] != self.n_var:
            raise Exception('Input dimension %s are not equal to n_var %s!' % (X.shape[1], self.n_var))
        if isinstance(return_values_of, str) and return_values_of == "auto":
            return_values_of = ["F", "CV"] if self.n_constr > 0 else ["F"]
        # a derivative ("d…") that `_evaluate` does not set by default switches gradients on (:100-107); 'hF' is the
        # reference's extension for Hessians (:110)
        gradient = any(v.startswith("d") for v in return_values_of if v not in self.evaluation_of)
        hessian = isinstance(return_values_of, list) and 'hF' in return_values_of
        out = dict.fromkeys(return_values_of)
        self._evaluate(X, out, *args, gradient=gradient, hessian=hessian, **kwargs)
        for key, val in out.items():  # pymoo `at_least2d`
            if isinstance(val, np.ndarray) and val.ndim == 1:
                out[key] = val[:, None]
        CV = np.zeros((X.shape[0], 1)) if self.n_constr == 0 else calc_constraint_violation(out["G"])  # :136-139
        if "CV" in return_values_of:
            out["CV"] = CV
        if "feasible" in return_values_of:
            out["feasible"] = CV <= 0
        for v in return_values_of:
            out.setdefault(v, None)
        if single:
            out = {k: (v if v is None else v[0, :]) for k, v in out.items()}
        if return_as_dictionary:
            return out
        picked = tuple(out[v] for v in return_values_of)
        return picked[0] if len(picked) == 1 else picked

    def evaluate_constraint(self, X):
        """Constraint values of continuous designs through the real problem (surrogate_problem.py:168-193)."""
        X = self.transformation.undo(X)
        if X.ndim == 1:
            return self.problem.evaluate_constraint(X)
        if X.ndim != 2:
            raise NotImplementedError
        G = np.array([self.problem.evaluate_constraint(x) for x in X])
        return None if None in G else G
