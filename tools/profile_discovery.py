"""cProfile of one batched ParetoDiscovery generation (bench --mode discovery shape)."""
import os, sys, cProfile, pstats
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autooed_b200 import GaussianProcess, Identity
from autooed_b200.problem import SimpleProblem
from autooed_b200.solver.pareto_discovery import pareto_discover
from autooed_b200.surrogate_problem import SurrogateProblem
d, m, N, pop = 10, 2, 100, 100
rng = np.random.default_rng(0)
X = rng.random((N, d)); f1 = X[:, 0]; gz = 1 + 9.0 / (d - 1) * X[:, 1:].sum(1)
Y = np.stack([f1, gz * (1 - np.sqrt(f1 / gz) - f1 / gz * np.sin(10 * np.pi * f1))], 1)
prob = SimpleProblem(d, m); gp = GaussianProcess(prob, nu=5); gp.fit(X, Y); acq = Identity(gp); acq.fit(X, Y)
sp = SurrogateProblem(prob, acq)
ev = lambda x, return_values_of: sp.evaluate(x, return_values_of=return_values_of)
xs = rng.random((pop, d)); bounds = np.stack([np.zeros(d), np.ones(d)]); origin = Y.min(axis=0)
np.random.seed(5); pareto_discover(xs, ev, bounds, None, 0.3, origin, 1e-2, 10)
pr = cProfile.Profile(); pr.enable()
np.random.seed(5); pareto_discover(xs, ev, bounds, None, 0.3, origin, 1e-2, 10)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(16)
