"""GPU parity of the Thompson-sampling acquisition against the reference-generated fixtures
(`tests/golden/ts_*.npz`) and the CPU oracle, plus the Uncertainty / Direct selections."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import ts_oracle as to

pytestmark = pytest.mark.gpu
FILES = sorted(f for f in os.listdir(GOLDEN) if f.startswith("ts_"))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def fitted(g, thetas=None):
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    gp = GaussianProcess(SimpleProblem(g["X"].shape[1], g["Y"].shape[1]), nu=int(g["nu"]))
    gp.fit_fixed(g["X"], g["Y"], g["gp_theta"] if thetas is None else thetas)
    return gp


@pytest.mark.parametrize("name", FILES)
def test_evaluate_reference_sample_on_gpu(name):
    """The reference's own sample (W, b, theta) evaluated by the CUDA kernel: value, gradient, Hessian."""
    from autooed_b200 import ThompsonSampling
    g = load(name)
    acq = ThompsonSampling(fitted(g))
    acq.thetas, acq.Ws, acq.bs, acq.sf2s, acq.fitted = list(g["theta"]), list(g["W"]), list(g["b"]), list(g["sf2"]), True
    F, dF, hF = acq.evaluate(g["Xs"], gradient=True, hessian=True)
    assert np.allclose(F, g["F"], rtol=1e-11, atol=1e-11)
    assert np.allclose(dF, g["dF"], rtol=1e-10, atol=1e-10)
    assert np.allclose(hF, g["hF"], rtol=1e-10, atol=1e-9)
    F2, none, none2 = acq.evaluate(g["Xs"])
    assert none is None and none2 is None and np.array_equal(F2, F)


@pytest.mark.parametrize("name", FILES)
def test_fit_draws_the_reference_sample_under_the_same_seed(name):
    """Same GP hyper-parameters + numpy seed 11 => the reference's spectral points / phases bit for bit and its weight
    sample up to the conditioning of the M x M system; evaluate then reproduces the stored outputs."""
    from autooed_b200 import ThompsonSampling
    g = load(name)
    acq = ThompsonSampling(fitted(g), n_spectral_pts=100)
    np.random.seed(int(g["seed"]))
    acq.fit(g["X"], g["Y"])
    assert np.array_equal(np.stack(acq.Ws), g["W"]) and np.array_equal(np.stack(acq.bs), g["b"])
    assert np.allclose(np.stack(acq.thetas), g["theta"], rtol=1e-6, atol=1e-8)
    F, dF, _ = acq.evaluate(g["Xs"], gradient=True)
    assert np.allclose(F, g["F"], rtol=1e-6, atol=1e-7) and np.allclose(dF, g["dF"], rtol=1e-5, atol=1e-6)


def test_large_population_vs_oracle_and_full_fit():
    """Device fit (own hyper-parameters) + a 50k-row population, d = 20, m = 3, against the oracle on the same sample."""
    from autooed_b200 import GaussianProcess, ThompsonSampling
    from autooed_b200.problem import SimpleProblem
    rng = np.random.default_rng(0)
    N, d, m = 150, 20, 3
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, m))) + 0.05 * rng.standard_normal((N, m))
    gp = GaussianProcess(SimpleProblem(d, m), nu=3)
    gp.fit(X, Y)
    acq = ThompsonSampling(gp)
    np.random.seed(1)
    acq.fit(X, Y)
    Xs = rng.random((50000, d))
    F, dF, _ = acq.evaluate(Xs, gradient=True)
    ys = gp.normalization.y_scaler
    oF, odF, _ = to.ts_evaluate(Xs[:3000], acq.thetas, acq.Ws, acq.bs, acq.sf2s, ys.mean_, ys.scale_, True, False)
    assert np.allclose(F[:3000], oF, rtol=1e-10, atol=1e-10) and np.allclose(dF[:3000], odF, rtol=1e-9, atol=1e-9)
    assert np.isfinite(F).all() and np.isfinite(dF).all()


@pytest.mark.parametrize("d", [40, 64, 100])
def test_more_than_32_variables(d):
    """33 .. 128 design variables: the wider template instances of the fit, posterior and sample kernels, end to end
    (fit on the device, Thompson sample value / gradient / Hessian against the oracle, host NSGA-II on top of it)."""
    from autooed_b200 import GaussianProcess, NSGA2, ThompsonSampling
    from autooed_b200.problem import SimpleProblem
    rng = np.random.default_rng(d)
    N, m = 60, 2
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, m)) / np.sqrt(d)) + 0.05 * rng.standard_normal((N, m))
    prob = SimpleProblem(d, m)
    gp = GaussianProcess(prob, nu=5)
    gp.fit(X, Y)
    acq = ThompsonSampling(gp)
    np.random.seed(1)
    acq.fit(X, Y)
    Xs = rng.random((300, d))
    F, dF, hF = acq.evaluate(Xs[:20], gradient=True, hessian=True)
    ys = gp.normalization.y_scaler
    oF, odF, ohF = to.ts_evaluate(Xs[:20], acq.thetas, acq.Ws, acq.bs, acq.sf2s, ys.mean_, ys.scale_, True, True)
    assert np.allclose(F, oF, rtol=1e-10, atol=1e-10) and np.allclose(dF, odF, rtol=1e-9, atol=1e-9)
    assert np.allclose(hF, ohF, rtol=1e-9, atol=1e-9)
    np.random.seed(2)
    Xc, Fc = NSGA2(prob, n_gen=5, pop_size=30).solve(X, Y, 4, acq)  # host loop: the device loop takes n_var <= 32
    assert Xc.shape == (30, d) and np.allclose(Fc, acq.evaluate(Xc)[0], rtol=1e-10, atol=1e-10)


def test_uncertainty_and_direct_selection():
    from autooed_b200 import Direct, GaussianProcess, Uncertainty
    from autooed_b200.problem import SimpleProblem
    rng = np.random.default_rng(2)
    X, Y = rng.random((40, 5)), rng.random((40, 2))
    gp = GaussianProcess(SimpleProblem(5, 2), nu=5)
    gp.fit(X, Y)
    Xc = rng.random((300, 5))
    picked = Uncertainty(gp).select(Xc, np.zeros((300, 2)), X, Y, 7)
    S = gp.evaluate(Xc, std=True)["S"]
    expect = Xc[np.argsort(np.prod(S, axis=1))[::-1][:7]]  # uncertainty.py:16-19
    assert np.array_equal(picked, expect)
    assert np.array_equal(Direct(gp).select(Xc[:4], np.zeros((4, 2)), X, Y, 4), Xc[:4])
    with pytest.raises(AssertionError):
        Direct(gp).select(Xc[:5], np.zeros((5, 2)), X, Y, 4)
