"""Per-phase cycle counts of lml_small_kernel (AOED_FIT_STAMP=1) — tuning aid."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["AOED_FIT_STAMP"] = "1"
from autooed_b200 import GaussianProcess
from autooed_b200.problem import SimpleProblem
for N, d in ((50, 6), (200, 6), (200, 20), (232, 10)):
    rng = np.random.default_rng(0)
    X = rng.random((N, d)); Yn = rng.standard_normal((N, 2))
    gp = GaussianProcess(SimpleProblem(d, 2), nu=5)
    gp._set_train(X, Yn)
    th = np.tile(np.concatenate([[0.0], np.zeros(d), [-4.0]]), (2, 1))
    gp.log_marginal_likelihood_batch([0, 1], th, True)
    gp.log_marginal_likelihood_batch([0, 1], th, True)
