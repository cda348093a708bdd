import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipy.optimize import minimize
from autooed_b200 import GaussianProcess, Identity
from autooed_b200.problem import SimpleProblem
from autooed_b200.solver import pareto_discovery as pdv
from autooed_b200.surrogate_problem import SurrogateProblem
from oracle import discovery_oracle as do, gp_oracle as go
d, m, N, pop = 10, 2, 100, 20
rng = np.random.default_rng(0)
X = rng.random((N, d)); f1 = X[:, 0]; gz = 1 + 9.0 / (d - 1) * X[:, 1:].sum(1)
Y = np.stack([f1, gz * (1 - np.sqrt(f1 / gz) - f1 / gz * np.sin(10 * np.pi * f1))], 1)
prob = SimpleProblem(d, m); gp = GaussianProcess(prob, nu=5); gp.fit(X, Y); acq = Identity(gp); acq.fit(X, Y)
sp = SurrogateProblem(prob, acq)
ev = lambda x, return_values_of: sp.evaluate(x, return_values_of=return_values_of)
xs = rng.random((100, d))[:pop]; bounds = np.stack([np.zeros(d), np.ones(d)]); origin = Y.min(axis=0)
ys = ev(xs, ['F']); no = np.minimum(origin, ys.min(0)); no = no - 1e-2 if (no != origin).any() else no
fs = ys - no; zs = np.array([pdv.reference_point(y, f, 0.3) for y, f in zip(ys, fs)])
xb = pdv.local_optimization_batched(xs, zs, ev, bounds)
orc = go.GPOracle(d, m, nu=5).fit(X, Y, thetas=np.stack([g.kernel_.theta for g in gp.gps]))
oev = do.oracle_eval_func(orc)
print("model agreement F:", np.abs(oev(xs, ['F']) - ys).max(), "dF:", np.abs(oev(xs, ['F', 'dF'])[1] - ev(xs, ['F', 'dF'])[1]).max())
for i in range(pop):
    xo = do.local_optimization(xs[i], ys[i], fs[i], oev, bounds, 0.3)
    xg = minimize(lambda x: pdv._distance_and_gradient(*ev(x, ['F', 'dF']), zs[i]), xs[i], method='L-BFGS-B', jac=True, bounds=bounds.T).x
    fo = np.linalg.norm(oev(xo, ['F']) - zs[i]); fb = np.linalg.norm(oev(xb[i], ['F']) - zs[i])
    print(i, "lockstep-vs-scipy(gpu)", np.abs(xb[i] - xg).max(), "gpu-vs-cpu", np.abs(xb[i] - xo).max(), "obj cpu", fo, "obj gpu", fb)
