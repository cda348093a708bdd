"""Exercises the non-dominant kernels once (fit at N=200, Hessian call, HVI selection, Pareto rank) so that one
`ncu --set full` pass can capture them (tools/gpu_profile_others.sh)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autooed_b200 import GaussianProcess, ExpectedImprovement, HypervolumeImprovement
from autooed_b200.problem import SimpleProblem
from autooed_b200.utils.pareto import pareto_rank
rng = np.random.default_rng(0)
N, d, m = 200, 6, 2
X = rng.random((N, d)); Y = np.sin(X @ rng.standard_normal((d, m))) + 0.05 * rng.standard_normal((N, m))
gp = GaussianProcess(SimpleProblem(d, m), nu=5, n_restarts=15, seed=0)
gp.fit(X, Y)
acq = ExpectedImprovement(gp); acq.fit(X, Y)
Xs = rng.random((1000, d))
F, dF, hF = acq.evaluate(Xs[:100], gradient=True, hessian=True)
F, _, _ = acq.evaluate(Xs)
HypervolumeImprovement(gp).select(Xs, F, X, Y, 10)
pareto_rank(np.vstack([F, Y]))
print("ok", gp.fit_info["n_batches"])
