"""One synchronous MOBO proposal step — the four calls of the reference's `MOBO._optimize`
(`autooed/mobo/mobo.py:98-114`): fit surrogate, fit acquisition, solve the surrogate problem, select the batch."""
from autooed_b200.utils.pareto import convert_minimization


class MOBO:
    def __init__(self, problem, surrogate_model, acquisition, solver, selection):
        self.problem = problem
        self.obj_type = getattr(problem, 'obj_type', None)
        self.surrogate_model, self.acquisition, self.solver, self.selection = surrogate_model, acquisition, solver, selection

    def optimize(self, X, Y, X_busy=None, batch_size=1):
        Y = convert_minimization(Y, self.obj_type)  # mobo.py:91
        self.surrogate_model.fit(X, Y)
        self.acquisition.fit(X, Y)
        X_candidate, Y_candidate = self.solver.solve(X, Y, batch_size, self.acquisition)
        return self.selection.select(X_candidate, Y_candidate, X, Y, batch_size)
