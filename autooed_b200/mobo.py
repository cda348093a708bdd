"""One MOBO proposal step — the reference's `MOBO.optimize / _optimize / _optimize_async / predict`
(`autooed/mobo/mobo.py:70-165`): fit surrogate, fit acquisition (or let the asynchronous strategy do both around the
designs still being evaluated), solve the surrogate problem, select the batch."""
from autooed_b200.utils.pareto import convert_minimization


class MOBO:
    def __init__(self, problem, surrogate_model, acquisition, solver, selection, async_strategy=None):
        self.problem = problem
        self.obj_type = getattr(problem, 'obj_type', None)
        self.surrogate_model, self.acquisition, self.solver, self.selection = surrogate_model, acquisition, solver, selection
        self.async_strategy = async_strategy

    def optimize(self, X, Y, X_busy=None, batch_size=1):
        Y = convert_minimization(Y, self.obj_type)  # mobo.py:91
        if self.async_strategy is None or X_busy is None:
            self.surrogate_model.fit(X, Y)
            self.acquisition.fit(X, Y)
            acquisition = self.acquisition
        else:  # mobo.py:116-129: the strategy may add believed data and / or swap in a penalised acquisition
            X, Y, acquisition = self.async_strategy.fit(X, Y, X_busy)
        X_candidate, Y_candidate = self.solver.solve(X, Y, batch_size, acquisition)
        return self.selection.select(X_candidate, Y_candidate, X, Y, batch_size)

    def predict(self, X, Y, X_next, fit=False):
        """Posterior mean (in the problem's own max / min sense) and standard deviation at X_next (mobo.py:131-165)."""
        Y = convert_minimization(Y, self.obj_type)
        if fit or self.async_strategy is not None:
            self.surrogate_model.fit(X, Y)
        mean, std = self.surrogate_model.predict(X_next, std=True)
        return convert_minimization(mean, self.obj_type), std
