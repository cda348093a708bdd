"""How far are the device LML / gradient and the CPU oracle from an 80-bit (numpy.longdouble) evaluation of the same
formulas on an ill-conditioned problem?  python tools/lml_accuracy.py [N d nu]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gp_oracle as go  # noqa: E402


def truth_lml(theta, X, y, nu):
    ld = np.longdouble
    K = go.kernel_matrix(theta, X, nu).astype(ld)
    n = len(y)
    K[np.diag_indices(n)] += ld(1e-10)
    L = np.zeros_like(K)
    for j in range(n):  # column Cholesky in extended precision
        s = K[j, j] - np.dot(L[j, :j], L[j, :j])
        L[j, j] = np.sqrt(s)
        if j + 1 < n:
            L[j + 1:, j] = (K[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    z = np.zeros(n, ld)
    for i in range(n):
        z[i] = (ld(y[i]) - np.dot(L[i, :i], z[:i])) / L[i, i]
    return float(-0.5 * np.dot(z, z) - np.log(np.diag(L)).sum() - 0.5 * n * np.log(2 * np.pi)), float(np.linalg.cond(K.astype(float)))


if __name__ == "__main__":
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    N, d, nu = (int(a) for a in (sys.argv[1:4] or [700, 3, 5]))
    os.environ["AOED_FIT_PATH"] = "blocked"
    rng = np.random.default_rng(7 * N + nu)
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, 1))) + 0.05 * rng.standard_normal((N, 1))
    Yn = (Y - Y.mean(0)) / Y.std(0)
    lo, hi = go.theta_bounds(d).T
    th = rng.uniform(0.4 * lo, 0.4 * hi, (3, d + 2))
    th[:, -1] = rng.uniform(-5.0, -2.0, 3)
    gp = GaussianProcess(SimpleProblem(d, 1), nu=nu)
    gp._set_train(X, Yn)
    lml, grad = gp.log_marginal_likelihood_batch([0, 0, 0], th, True)
    for b in range(3):
        t, cond = truth_lml(th[b], X, Yn[:, 0], nu)
        o, og = go.lml_and_grad(th[b], X, Yn[:, 0], nu)
        print(f"cond {cond:.2e} truth {t:.10e}  gpu err {abs(lml[b] - t) / abs(t):.2e}  oracle err {abs(o - t) / abs(t):.2e}  "
              f"grad gpu-vs-oracle {np.abs(grad[b] - og).max() / np.abs(og).max():.2e}", flush=True)
