import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autooed_b200 import GaussianProcess
from autooed_b200.problem import SimpleProblem
from oracle import gp_oracle as go
N, d, m = 2000, 20, 3
rng = np.random.default_rng(0)
X = rng.random((N, d)); Y = np.sin(X @ rng.standard_normal((d, m))) + 0.02 * rng.standard_normal((N, m))
Yn = (Y - Y.mean(0)) / Y.std(0)
gp = GaussianProcess(SimpleProblem(d, m), nu=5); gp._set_train(X, Yn)
lo, hi = go.theta_bounds(d).T
order = [48, 1, 2, 3, 6, 9, 12, 16, 24, 32, 48, 12]
for B in order:
    obj = np.arange(B) % m; th = rng.uniform(lo, hi, (B, d + 2)) * 0.3
    gp.log_marginal_likelihood_batch(obj, th, True)
    t0 = time.perf_counter()
    for _ in range(3): gp.log_marginal_likelihood_batch(obj, th, True)
    print(B, round(1e3 * (time.perf_counter() - t0) / 3, 2), "ms", flush=True)
