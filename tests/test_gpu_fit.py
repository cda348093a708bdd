"""GPU parity of the hyper-parameter fit (through the C ABI): batched log-marginal-likelihood + gradient against the
reference-generated golden vectors (sklearn `log_marginal_likelihood`, _gpr.py:541-655 via gp.py:139) and the CPU
oracle, for both device paths (one-CTA-per-problem shared-memory kernel, blocked cuSOLVER path), and the full
L-BFGS-B fit against the reference's fitted theta / LML."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_files
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu
FILES = golden_files()


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def model_for(g, **kw):
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    prob = SimpleProblem(g["X"].shape[1], g["Y"].shape[1], g["xl"], g["xu"])
    return GaussianProcess(prob, nu=int(g["nu"]), **kw)


def grad_tol(nu):
    # the gradient contracts (alpha alpha^T - K^-1) with dK: rounding scales with |K^-1| (same allowance as the oracle's pin)
    return 1e-6 if nu in (5, -1) else 1e-8


def check_lml_batch(name, counter):
    g = load(name)
    nu = int(g["nu"])
    gp = model_for(g)
    gp._set_train(g["Xn"], g["Yn"])
    m = g["Y"].shape[1]
    obj = np.repeat(np.arange(m), 3)
    thetas = g["lml_thetas"].reshape(m * 3, -1)
    lml, grad = gp.log_marginal_likelihood_batch(obj, thetas, True)
    ref_v, ref_g = g["lml_vals"].reshape(-1), g["lml_grads"].reshape(m * 3, -1)
    assert np.all(np.abs(lml - ref_v) <= 1e-8 * np.maximum(1.0, np.abs(ref_v)))
    for b in range(m * 3):
        assert np.abs(grad[b] - ref_g[b]).max() <= grad_tol(nu) * max(1.0, np.abs(ref_g[b]).max()), (b, grad[b], ref_g[b])
    # value-only call gives the same values
    lml2, none = gp.log_marginal_likelihood_batch(obj, thetas, False)
    assert none is None and np.array_equal(lml, lml2)
    assert gp.fit_counters()[counter] == 2 * m * 3


@pytest.mark.parametrize("path", ["small", "blocked"])
@pytest.mark.parametrize("name", FILES)
def test_lml_batch_vs_reference(name, path, monkeypatch):
    """All (objective x probe theta) problems of a fixture in ONE batched call, on the one-CTA-per-problem kernel
    (blocked DMMA version) and on the multi-stream path for large N."""
    if path == "blocked":
        monkeypatch.setenv("AOED_FIT_PATH", "blocked")
    check_lml_batch(name, "blocked" if path == "blocked" else "small")


def test_lml_batch_rank1_kernel_vs_reference():
    """The rank-1 predecessor of the one-CTA kernel still serves 208 < N <= 232; the choice is latched per process, so
    the fixtures are replayed in a child process with AOED_FIT_SMALL=v1."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = ("import os, sys; os.environ['AOED_FIT_SMALL'] = 'v1'; sys.path[:0] = [%r, %r]; import test_gpu_fit as t; "
            "[t.check_lml_batch(n, 'small') for n in t.FILES]; print('ok')" % (here, os.path.dirname(here)))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stderr[-2000:]


@pytest.mark.parametrize("nu", [1, 3, 5, -1])
@pytest.mark.parametrize("shape", [(1, 2, 1), (2, 3, 2), (33, 5, 2), (232, 9, 2), (233, 9, 1), (600, 20, 3), (100, 40, 2),
                                   (190, 64, 2), (230, 64, 4), (300, 50, 1), (120, 100, 2), (260, 128, 1)])
def test_lml_batch_vs_oracle(nu, shape):
    """Edge sizes (N = 1, the shared-memory limit 232 / 233, a multi-stream blocked problem) and a non-PD theta."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    N, d, m = shape
    rng = np.random.default_rng(11 * N + d + nu)
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, m))) + 0.05 * rng.standard_normal((N, m))
    Yn = (Y - Y.mean(0)) / np.where(Y.std(0) > 0, Y.std(0), 1.0)
    gp = GaussianProcess(SimpleProblem(d, m), nu=nu)
    gp._set_train(X, Yn)
    lo, hi = go.theta_bounds(d).T
    n_prob = 2 * m + 1
    obj = np.arange(n_prob) % m
    thetas = rng.uniform(lo, hi, (n_prob, d + 2))
    thetas[:, -1] = rng.uniform(-6, -2, n_prob)
    lml, grad = gp.log_marginal_likelihood_batch(obj, thetas, True)
    for b in range(n_prob):
        rv, rg = go.lml_and_grad(thetas[b], X, Yn[:, obj[b]], nu)
        if not np.isfinite(rv):
            assert lml[b] == -np.inf and not grad[b].any()
            continue
        assert abs(lml[b] - rv) <= 1e-7 * max(1.0, abs(rv)), (b, lml[b], rv)
        assert np.abs(grad[b] - rg).max() <= 1e-5 * max(1.0, np.abs(rg).max()), (b, grad[b], rg)


def test_non_pd_reports_minus_inf():
    """Duplicate training rows with a smooth kernel: K is singular to working precision (sklearn: -inf, zero gradient)."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    rng = np.random.default_rng(0)
    X = np.repeat(rng.random((10, 3)), 4, axis=0)
    Yn = rng.standard_normal((40, 1))
    th = np.array([[3.4, 3.4, 3.4, 3.4, -6.0]])
    for nu in (5, -1):
        rv, rg = go.lml_and_grad(th[0], X, Yn[:, 0], nu)
        gp = GaussianProcess(SimpleProblem(3, 1), nu=nu)
        gp._set_train(X, Yn)
        lml, grad = gp.log_marginal_likelihood_batch([0], th, True)
        if not np.isfinite(rv):
            assert lml[0] == -np.inf and not grad[0].any()


@pytest.mark.parametrize("name", FILES)
def test_fit_reaches_reference_optimum(name):
    """fit(X, Y): same optimiser (scipy L-BFGS-B, gp.py:94-101) driven by the device LML.  The optimiser path may
    differ in low bits (scipy 1.18 C port vs the pinned Fortran build), so parity is on the optimum:
    LML at least as good as the reference's up to 1e-6 relative, and the same theta where the optimum is sharp."""
    g = load(name)
    nu = int(g["nu"])
    gp = model_for(g)
    gp.fit(g["X"], g["Y"])
    for o in range(g["Y"].shape[1]):
        got = gp.gps[o].log_marginal_likelihood_value_
        ref = g["lml_star"][o]
        # RBF on noise-free data: cond(K) ~ 1e13, the LML gradient is rounding noise near the optimum and the CPU
        # oracle itself (same scipy, same formulas) stops 2.4e-5 below the stored sklearn run on zdt1/rbf
        tol = 1e-4 if nu == -1 else 1e-6
        assert got >= ref - tol * max(1.0, abs(ref)), (o, got, ref)
        _, orc_lml, _ = go.fit_theta(g["Xn"], g["Yn"][:, o], nu)
        assert abs(got - orc_lml) <= tol * max(1.0, abs(orc_lml)), (o, got, orc_lml)
        # the oracle at the GPU's theta reproduces the GPU's LML (the device value is not self-consistent only)
        rv, _ = go.lml_and_grad(gp.gps[o].kernel_.theta, g["Xn"], g["Yn"][:, o], int(g["nu"]), want_grad=False)
        assert abs(got - rv) <= 1e-8 * max(1.0, abs(rv))
    assert gp.fit_info["n_problems"] == g["Y"].shape[1]


def test_fit_matches_oracle_trajectory():
    """Well-conditioned noisy data: the device-driven L-BFGS-B lands on the oracle's theta (same optimiser, same starts)."""
    g = load("noisy_d4_n50_nu1.npz")
    gp = model_for(g)
    gp.fit(g["X"], g["Y"])
    orc = go.GPOracle(g["X"].shape[1], g["Y"].shape[1], g["xl"], g["xu"], int(g["nu"])).fit(g["X"], g["Y"])
    for o, st in enumerate(orc.states):
        assert np.allclose(gp.gps[o].kernel_.theta, st.theta, rtol=1e-4, atol=1e-4)
        assert np.allclose(gp.gps[o].kernel_.theta, g["theta"][o], rtol=1e-3, atol=1e-3)


def test_restarts_keep_best_and_batch_together():
    g = load("zdt1_d6_n40_nu5.npz")
    base = model_for(g)
    base.fit(g["X"], g["Y"])
    multi = model_for(g, n_restarts=5, seed=3)
    multi.fit(g["X"], g["Y"])
    m = g["Y"].shape[1]
    assert multi.fit_info["n_problems"] == m * 6
    # lock-step: far fewer device batches than LML evaluations
    assert multi.fit_info["n_batches"] < multi.fit_info["n_lml_evals"] / 2
    for o in range(m):
        assert multi.gps[o].log_marginal_likelihood_value_ >= base.gps[o].log_marginal_likelihood_value_ - 1e-9
    # sklearn attribute surface after fit
    assert multi.gps[0].L_.shape == (40, 40) and multi.gps[0].alpha_.shape == (40,)


@pytest.mark.parametrize("n_restarts", [0, 3])
def test_lockstep_driver_equals_threaded_driver(n_restarts, monkeypatch):
    """The fit drives scipy's L-BFGS-B routine by reverse communication in one thread (lbfgsb_lockstep.py);
    AOED_FIT_DRIVER=threads runs `scipy.optimize.minimize` in one thread per problem instead.  Same routine, same
    arguments, same device evaluations => identical hyper-parameters and evaluation counts."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    rng = np.random.default_rng(4)
    X = rng.random((70, 5))
    Y = np.stack([np.sin(3 * X[:, 0]) + X[:, 1] ** 2, np.cos(2 * X[:, 2]) * X[:, 3], X.sum(1)], 1)
    out = {}
    for driver in ("lockstep", "threads"):
        monkeypatch.setenv("AOED_FIT_DRIVER", driver)
        gp = GaussianProcess(SimpleProblem(5, 3), nu=5, n_restarts=n_restarts, seed=1)
        gp.fit(X, Y)
        assert gp.fit_info["driver"] == driver
        out[driver] = (np.stack([g.kernel_.theta for g in gp.gps]), gp.fit_info["nfev"], gp.fit_info["n_lml_evals"])
    assert np.array_equal(out["lockstep"][0], out["threads"][0])
    assert out["lockstep"][1] == out["threads"][1] and out["lockstep"][2] == out["threads"][2]
