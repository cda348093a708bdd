"""Smallest shapes that still reach every synchronisation protocol of the library (TMA + mbarrier ring of trmm_tma_kernel
and tgemm_kernel, the one-CTA shared-memory Cholesky kernels, the NSGA-II survival kernels, the HVI round kernel) — the
workload of the compute-sanitizer runs:
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import autooed_b200 as ab  # noqa: E402
from autooed_b200.problem import SimpleProblem  # noqa: E402
from autooed_b200.selection.hvi import hvi_select_indices  # noqa: E402

rng = np.random.default_rng(0)
which = sys.argv[1:] or ["eval", "fit_small", "fit_blocked", "nsga2", "hvi"]
d, m = 4, 2
if "eval" in which:  # N = 130: two row tiles, persistent TMA kernel LOWER + UPPER, gradient, Hessian
    X = rng.random((130, d))
    Y = np.sin(X @ rng.standard_normal((d, m)))
    th = np.tile(np.concatenate([[0.0], np.full(d, -0.5), [-3.0]]), (m, 1))
    gp = ab.GaussianProcess(SimpleProblem(d, m), nu=5)
    gp.fit_fixed(X, Y, th)
    out = gp.evaluate(rng.random((300, d)), std=True, gradient=True)
    out2 = gp.evaluate(rng.random((3, d)), std=True, gradient=True, hessian=True)
    print("eval ok", float(out["S"].mean()), float(np.abs(out2["hS"]).max()))
if "fit_small" in which:  # N = 64: one-CTA blocked kernel, 4 problems
    X = rng.random((64, d))
    Y = np.sin(X @ rng.standard_normal((d, m)))
    gp = ab.GaussianProcess(SimpleProblem(d, m), nu=3, n_restarts=1, seed=0)
    gp.fit(X, Y)
    print("fit_small ok", gp.fit_info["n_lml_evals"])
if "fit_blocked" in which:  # N = 300: three tile rows of the blocked factorisation (tgemm NT / NN, potrf_diag), 3 problems
    os.environ["AOED_FIT_PATH"] = "blocked"
    X = rng.random((300, d))
    Yn = rng.standard_normal((300, m))
    gp = ab.GaussianProcess(SimpleProblem(d, m), nu=5)
    gp._set_train(X, Yn)
    th = np.tile(np.concatenate([[0.0], np.full(d, -1.0), [-3.0]]), (3, 1))
    lml, grad = gp.log_marginal_likelihood_batch([0, 1, 0], th, True)
    print("fit_blocked ok", lml)
    os.environ.pop("AOED_FIT_PATH")
if "nsga2" in which:  # pop 100 x 3 generations on the device
    X = rng.random((40, d))
    Y = np.stack([X[:, 0], 1 + X[:, 1:].sum(1) - np.sqrt(X[:, 0])], 1)
    prob = SimpleProblem(d, m)
    gp = ab.GaussianProcess(prob, nu=1)
    gp.fit(X, Y)
    acq = ab.Identity(gp)
    acq.fit(X, Y)
    solver = ab.init_solver('nsga2', {'n_gen': 3, 'pop_size': 100}, prob)
    np.random.seed(0)
    Xc, Yc = solver.solve(X, Y, 5, acq)
    print("nsga2 ok", Xc.shape)
if "hvi" in which:
    Y = rng.random((30, 3)) + 0.2
    Yc = rng.random((700, 3))
    np.random.seed(0)
    print("hvi ok", hvi_select_indices(Yc, Y, 5))
