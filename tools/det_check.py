"""Bitwise reproducibility of the one-CTA LML kernel across repetitions and batch compositions."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autooed_b200 import GaussianProcess
from autooed_b200.problem import SimpleProblem
rng = np.random.default_rng(4)
X = rng.random((70, 5))
Y = np.stack([np.sin(3 * X[:, 0]) + X[:, 1] ** 2, np.cos(2 * X[:, 2]) * X[:, 3], X.sum(1)], 1)
Yn = (Y - Y.mean(0)) / Y.std(0)
gp = GaussianProcess(SimpleProblem(5, 3), nu=5)
gp._set_train(X, Yn)
th = rng.uniform(-2, 2, (6, 7)); th[:, -1] = -4
obj = np.array([0, 1, 2, 2, 1, 0])
ref_l, ref_g = gp.log_marginal_likelihood_batch(obj, th, True)
bad = 0
for rep in range(200):
    l, g = gp.log_marginal_likelihood_batch(obj, th, True)
    if not (np.array_equal(l, ref_l) and np.array_equal(g, ref_g)):
        bad += 1
print("repeat mismatches", bad)
# single problems / permuted batches
bad2 = 0
for rep in range(50):
    perm = rng.permutation(6)
    l, g = gp.log_marginal_likelihood_batch(obj[perm], th[perm], True)
    if not (np.array_equal(l, ref_l[perm]) and np.array_equal(g, ref_g[perm])):
        bad2 += 1
        if bad2 < 3:
            print("perm diff", np.abs(l - ref_l[perm]).max(), np.abs(g - ref_g[perm]).max())
for b in range(6):
    l, g = gp.log_marginal_likelihood_batch(obj[b:b + 1], th[b:b + 1], True)
    if not (l[0] == ref_l[b] and np.array_equal(g[0], ref_g[b])):
        bad2 += 1
        print("single diff", b, abs(l[0] - ref_l[b]), np.abs(g[0] - ref_g[b]).max())
print("composition mismatches", bad2)
