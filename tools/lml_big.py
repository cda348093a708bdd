"""One blocked-path LML + gradient batch at c4 size (N=2000, d=20): used under ncu to see where the time goes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autooed_b200 import GaussianProcess
from autooed_b200.problem import SimpleProblem
N, d, m = 2000, 20, 3
rng = np.random.default_rng(0)
X = rng.random((N, d)); Yn = rng.standard_normal((N, m))
gp = GaussianProcess(SimpleProblem(d, m), nu=5)
gp._set_train(X, Yn)
th = np.tile(np.concatenate([[0.0], np.zeros(d), [-4.0]]), (3, 1))
gp.log_marginal_likelihood_batch([0, 1, 2], th, True)
t = time.perf_counter()
gp.log_marginal_likelihood_batch([0, 1, 2], th, True)
print("batch of 3:", time.perf_counter() - t)
