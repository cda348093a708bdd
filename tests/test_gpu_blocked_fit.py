"""GPU parity of the blocked factorisation (autooed_b200/csrc/blocked.cu: left-looking tile Cholesky, recursive-doubling
triangular inverse, K^-1 on the FP64 tensor pipe) that serves log_marginal_likelihood for n_train > ~230 and the final
factorisation of every fit — against the CPU oracle (sklearn's formulas, _gpr.py:541-655 / :349-367) at sizes that
exercise 2 ... 9 tile rows, ragged last tiles, several problems per launch and several groups per call."""
import numpy as np
import pytest

from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu
EPS = np.finfo(float).eps


def data(N, d, m, seed, noise=0.05):
    rng = np.random.default_rng(seed)
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, m))) + noise * rng.standard_normal((N, m))
    Yn = (Y - Y.mean(0)) / Y.std(0)
    return rng, X, Y, Yn


def thetas_for(rng, d, n, c2_lo=-5.0):
    lo, hi = go.theta_bounds(d).T
    th = rng.uniform(0.4 * lo, 0.4 * hi, (n, d + 2))
    th[:, -1] = rng.uniform(c2_lo, -2.0, n)
    return th


@pytest.mark.parametrize("nu", [1, 5])
@pytest.mark.parametrize("shape", [(129, 4, 2), (256, 6, 1), (300, 7, 3), (513, 10, 2), (640, 12, 2), (700, 3, 1), (1100, 8, 2)])
def test_blocked_lml_and_gradient_vs_oracle(shape, nu, monkeypatch):
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    monkeypatch.setenv("AOED_FIT_PATH", "blocked")
    N, d, m = shape
    rng, X, Y, Yn = data(N, d, m, 7 * N + nu)
    gp = GaussianProcess(SimpleProblem(d, m), nu=nu)
    gp._set_train(X, Yn)
    n_prob = 2 * m + 1
    obj = np.arange(n_prob) % m
    th = thetas_for(rng, d, n_prob)
    lml, grad = gp.log_marginal_likelihood_batch(obj, th, True)
    lml2, _ = gp.log_marginal_likelihood_batch(obj, th, False)
    assert np.array_equal(lml, lml2)
    for b in range(n_prob):
        rv, rg = go.lml_and_grad(th[b], X, Yn[:, obj[b]], nu)
        assert np.isfinite(rv)
        # both sides solve with a Cholesky factor of K: agreement is bounded by cond(K) eps (1e-8 while cond(K) < 4e5)
        cond = np.linalg.cond(go.kernel_matrix(th[b], X, nu) + 1e-10 * np.eye(N))
        tol = max(1e-8, 0.1 * cond * EPS)
        assert abs(lml[b] - rv) <= tol * max(1.0, abs(rv)), (b, lml[b], rv, cond)
        assert np.abs(grad[b] - rg).max() <= 100 * tol * max(1.0, np.abs(rg).max()), (b, grad[b], rg, cond)
    assert gp.fit_counters()["blocked"] == 2 * n_prob


@pytest.mark.parametrize("shape", [(40, 3, 2), (200, 6, 2), (385, 5, 3), (900, 9, 2)])
def test_final_factorisation_state_vs_numpy(shape):
    """aoed_gp_set_theta: L_, alpha_, log_marginal_likelihood_value_ against numpy's Cholesky of the oracle's K."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    N, d, m = shape
    rng, X, Y, Yn = data(N, d, m, 3 * N)
    th = thetas_for(rng, d, m, c2_lo=-4.0)
    gp = GaussianProcess(SimpleProblem(d, m), nu=3)
    gp.fit_fixed(X, Y, th)
    Xn, Yn = gp._Xn, gp._Yn
    for o in range(m):
        K = go.kernel_matrix(th[o], Xn, 3) + 1e-10 * np.eye(N)
        L = np.linalg.cholesky(K)
        alpha = np.linalg.solve(K, Yn[:, o])
        cond = np.linalg.cond(K)
        tol = max(1e-10, 50 * cond * np.finfo(float).eps)
        got_L, got_a = gp.gps[o].L_, gp.gps[o].alpha_
        assert np.array_equal(got_L, np.tril(got_L))
        assert np.abs(got_L - L).max() <= tol * np.abs(L).max()
        assert np.abs(got_a - alpha).max() <= tol * np.abs(alpha).max()
        rv, _ = go.lml_and_grad(th[o], Xn, Yn[:, o], 3, want_grad=False)
        assert abs(gp.gps[o].log_marginal_likelihood_value_ - rv) <= 1e-9 * max(1.0, abs(rv))


def test_blocked_groups_when_scratch_is_small(monkeypatch):
    """With a scratch budget of one problem the batch is served in groups of one: same numbers as side by side."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    N, d, m = 300, 5, 2
    rng, X, Y, Yn = data(N, d, m, 5)
    th = thetas_for(rng, d, 5)
    obj = np.arange(5) % m
    out = []
    for gb in ("16", "0.005"):  # 0.005 GB = 5.4 MB: one 384 x 384 problem needs 4.7 MB
        monkeypatch.setenv("AOED_FIT_SCRATCH_GB", gb)
        gp = GaussianProcess(SimpleProblem(d, m), nu=5)
        gp._set_train(X, Yn)
        out.append(gp.log_marginal_likelihood_batch(obj, th, True))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_blocked_non_pd_problem_does_not_disturb_its_neighbours(monkeypatch):
    """One singular problem in the batch (duplicate rows, smooth kernel, huge length scales): -inf and a zero gradient for
    it (_gpr.py:589-593), untouched values for the others."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    monkeypatch.setenv("AOED_FIT_PATH", "blocked")
    rng = np.random.default_rng(0)
    X = np.repeat(rng.random((75, 3)), 4, axis=0)  # N = 300, every row four times
    Yn = rng.standard_normal((300, 1))
    bad = np.array([3.4, 3.4, 3.4, 3.4, -6.0])
    good = np.array([0.0, -2.0, -2.0, -2.0, -1.0])  # c2 = e^-1 on the diagonal does not help either: K has exact duplicates
    gp = GaussianProcess(SimpleProblem(3, 1), nu=-1)
    gp._set_train(X, Yn)
    lml, grad = gp.log_marginal_likelihood_batch([0, 0, 0], np.stack([good, bad, good]), True)
    for b, th in enumerate((good, bad, good)):
        rv, rg = go.lml_and_grad(th, X, Yn[:, 0], -1)
        if not np.isfinite(rv):
            assert lml[b] == -np.inf
        if np.isfinite(lml[b]):  # K + 1e-10 I with exact duplicate rows: cond ~ 1e10 / 1e-10, only the leading digits are defined
            assert abs(lml[b] - rv) <= 1e-3 * max(1.0, abs(rv))
        else:
            assert lml[b] == -np.inf and not grad[b].any()
    assert np.isfinite(lml[0])
    assert lml[0] == lml[2] and np.array_equal(grad[0], grad[2])
