"""Blocked-path LML + gradient batches at mid sizes (N = 300 / 500, the usual MOBO regime after a few dozen iterations):
timing, and under ncu the launch list of one batch."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autooed_b200 import GaussianProcess
from autooed_b200.problem import SimpleProblem
N, d, m = int(sys.argv[1]), 8, 2
rng = np.random.default_rng(0)
X = rng.random((N, d)); Yn = rng.standard_normal((N, m))
gp = GaussianProcess(SimpleProblem(d, m), nu=5)
gp._set_train(X, Yn)
th = np.tile(np.concatenate([[0.0], np.zeros(d), [-4.0]]), (m, 1))
for _ in range(3):
    gp.log_marginal_likelihood_batch(list(range(m)), th, True)
t = time.perf_counter()
for _ in range(10):
    gp.log_marginal_likelihood_batch(list(range(m)), th, True)
print(f"N {N}: batch of {m}: {(time.perf_counter() - t) / 10 * 1e3:.3f} ms")
