"""Solver-facing adapter — same call surface as the reference's `autooed/mobo/surrogate_problem.py:11-193`
(`SurrogateProblem.evaluate / evaluate_constraint`), i.e. the pymoo 0.4.1 `Problem` protocol that NSGA-II / MOEA-D /
ParetoDiscovery drive.  The acquisition it wraps evaluates on the GPU; this class is host logic only (flag decoding,
cheap-constraint loop, constraint violation) and does not need pymoo to be importable.  INTEGRATION.md shows how a
maintainer re-bases it on `pymoo.model.problem.Problem`."""
import numpy as np


def calc_constraint_violation(G):
    """pymoo 0.4.1 `Problem.calc_constraint_violation`: sum of the positive parts per row, as a column."""
    if G is None:
        return None
    if G.shape[1] == 0:
        return np.zeros(G.shape[0])[:, None]
    return np.sum(G * (G > 0).astype(float), axis=1)[:, None]


class SurrogateProblem:

    def __init__(self, problem, acquisition):
        self.problem = problem
        self.transformation = problem.transformation
        self.acquisition = acquisition
        self.n_var, self.n_obj = problem.n_var, problem.n_obj
        self.n_constr = getattr(problem, 'n_constr', 0)
        self.xl, self.xu = np.asarray(problem.xl, dtype=float), np.asarray(problem.xu, dtype=float)
        self.callback = None
        # what `_evaluate` sets by itself (pymoo's `evaluation_of="auto"`): F, and G with constraints
        self.evaluation_of = ['F'] + (['G'] if self.n_constr > 0 else [])

    def _evaluate(self, X, out, *args, gradient, hessian, **kwargs):
        # surrogate_problem.py:32-52
        out['F'], out['dF'], out['hF'] = self.acquisition.evaluate(X, dtype='continuous', gradient=gradient, hessian=hessian)
        X_raw = self.transformation.undo(X)
        out['G'] = np.array([self.problem.evaluate_constraint(x_raw) for x_raw in X_raw])

    def evaluate(self, X, *args, return_values_of="auto", return_as_dictionary=False, **kwargs):
        # surrogate_problem.py:54-166
        assert self.acquisition.fitted, 'Acquisition function is not fitted before surrogate problem evaluation'
        if self.callback is not None:
            self.callback(X)
        only_single_value = len(np.shape(X)) == 1
        X = np.atleast_2d(X)
        if X.shape[1] != self.n_var:
            raise Exception('Input dimension %s are not equal to n_var %s!' % (X.shape[1], self.n_var))
        if type(return_values_of) == str and return_values_of == "auto":
            return_values_of = ["F"]
            if self.n_constr > 0:
                return_values_of.append("CV")
        values_not_set = [val for val in return_values_of if val not in self.evaluation_of]
        gradient = len([val for val in values_not_set if val.startswith("d")]) > 0
        hessian = (type(return_values_of) == list and 'hF' in return_values_of)
        out = {val: None for val in return_values_of}
        self._evaluate(X, out, *args, gradient=gradient, hessian=hessian, **kwargs)
        for key in list(out.keys()):  # pymoo `at_least2d`
            if out[key] is not None and isinstance(out[key], np.ndarray) and out[key].ndim == 1:
                out[key] = out[key][:, None]
        if self.n_constr == 0:
            CV = np.zeros([X.shape[0], 1])
        else:
            CV = calc_constraint_violation(out["G"])
        if "CV" in return_values_of:
            out["CV"] = CV
        if "feasible" in return_values_of:
            out["feasible"] = (CV <= 0)
        for val in return_values_of:
            if val not in out:
                out[val] = None
        if only_single_value:
            for key in out.keys():
                if out[key] is not None:
                    out[key] = out[key][0, :]
        if return_as_dictionary:
            return out
        if len(return_values_of) == 1:
            return out[return_values_of[0]]
        return tuple([out[val] for val in return_values_of])

    def evaluate_constraint(self, X):
        # surrogate_problem.py:168-193
        X = self.transformation.undo(X)
        if X.ndim == 1:
            return self.problem.evaluate_constraint(X)
        elif X.ndim == 2:
            G = np.array([self.problem.evaluate_constraint(x) for x in X])
            if None in G:
                return None
            return G
        raise NotImplementedError
