"""Randomised parity sweep (not part of the test suite): final factorisation, posterior and batched LML + gradient against
the CPU oracle over random (N, d, m, nu), straddling the tile boundaries of the blocked factorisation."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autooed_b200 import GaussianProcess
from autooed_b200.problem import SimpleProblem
from oracle import gp_oracle as go
EPS = np.finfo(float).eps
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
worst = {"L": 0, "alpha": 0, "F": 0, "S": 0, "lml": 0, "grad": 0}
t0 = time.time()
for case in range(n_cases):
    N = int(rng.choice([rng.integers(1, 40), rng.integers(100, 140), rng.integers(120, 132), rng.integers(250, 262), rng.integers(380, 390),
                        rng.integers(200, 700), rng.integers(500, 520), rng.integers(700, 1100)]))
    d = int(rng.choice([1, 2, 3, 4, 5, 8, 9, 12, 16, 17, 20, 24, 28, 29, 33, 40, 48, 64, 70, 100, 128]))
    m = int(rng.integers(1, 4))
    nu = int(rng.choice([1, 3, 5, -1]))
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, m))) + 0.1 * rng.standard_normal((N, m))
    lo, hi = go.theta_bounds(d).T
    th = rng.uniform(0.3 * lo, 0.3 * hi, (m, d + 2)); th[:, -1] = rng.uniform(-4, -1.5, m)
    gp = GaussianProcess(SimpleProblem(d, m), nu=nu)
    gp.fit_fixed(X, Y, th)
    Xn, Yn = gp._Xn, gp._Yn
    orc = go.GPOracle(d, m, nu=nu).fit(X, Y, thetas=th)
    Xs = rng.random((37, d))
    out = gp.evaluate(Xs, std=True, gradient=True)
    ref = orc.evaluate(Xs, std=True, gradient=True, var_form="chol")
    for o in range(m):
        K = go.kernel_matrix(th[o], Xn, nu) + 1e-10 * np.eye(N)
        cond = np.linalg.cond(K)
        tol = max(1e-10, 50 * cond * EPS)
        L = np.linalg.cholesky(K)
        eL = np.abs(gp.gps[o].L_ - L).max() / np.abs(L).max() / tol
        ea = np.abs(gp.gps[o].alpha_ - np.linalg.solve(K, Yn[:, o])).max() / max(np.abs(gp.gps[o].alpha_).max(), 1e-300) / tol
        worst["L"], worst["alpha"] = max(worst["L"], eL), max(worst["alpha"], ea)
        assert eL < 1 and ea < 1, (case, N, d, m, nu, eL, ea, cond)
    tolF = max(1e-9, 100 * cond * EPS)
    eF = np.abs(out["F"] - ref["F"]).max() / max(np.abs(ref["F"]).max(), 1e-300) / tolF
    prior = np.exp(th[:, 0]) + np.exp(th[:, -1])
    eS = (np.abs((out["S"] / orc.norm.y_scale) ** 2 - (ref["S"] / orc.norm.y_scale) ** 2) / prior).max() / max(1e-10, tolF)
    worst["F"], worst["S"] = max(worst["F"], eF), max(worst["S"], eS)
    assert eF < 1 and eS < 1, (case, N, d, m, nu, eF, eS)
    B = 2 * m + 1
    obj = np.arange(B) % m
    thb = rng.uniform(0.3 * lo, 0.3 * hi, (B, d + 2)); thb[:, -1] = rng.uniform(-4, -1.5, B)
    lml, grad = gp.log_marginal_likelihood_batch(obj, thb, True)
    for b in range(B):
        rv, rg = go.lml_and_grad(thb[b], Xn, Yn[:, obj[b]], nu)
        condb = np.linalg.cond(go.kernel_matrix(thb[b], Xn, nu) + 1e-10 * np.eye(N))
        tl = max(1e-8, condb * EPS)  # both arms carry O(cond eps) in log|K| and the quadratic form
        el = abs(lml[b] - rv) / max(1.0, abs(rv)) / tl
        eg = np.abs(grad[b] - rg).max() / max(1.0, np.abs(rg).max()) / (100 * tl)
        worst["lml"], worst["grad"] = max(worst["lml"], el), max(worst["grad"], eg)
        assert el < 1 and eg < 1, (case, N, d, m, nu, b, el, eg, condb)
print(f"{n_cases} cases ok in {time.time() - t0:.0f}s; worst error / tolerance:", {k: round(v, 4) for k, v in worst.items()})
