"""GPU parity of the posterior / acquisition path (through the C ABI) against the reference-generated golden
fixtures and the CPU oracle.  Tolerances: F, dF <= 1e-8 relative (north star), widened only by the measured
conditioning of K where the reference itself loses digits."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_files
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu
FILES = golden_files()
EPS = np.finfo(float).eps


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def build(g, **kw):
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    prob = SimpleProblem(g["X"].shape[1], g["Y"].shape[1], g["xl"], g["xu"])
    gp = GaussianProcess(prob, nu=int(g["nu"]), **kw)
    gp.fit_fixed(g["X"], g["Y"], g["theta"])
    return gp


def oracle_from(g):
    orc = go.GPOracle(g["X"].shape[1], g["Y"].shape[1], g["xl"], g["xu"], int(g["nu"]))
    orc.fit(g["X"], g["Y"], thetas=g["theta"])
    return orc


def cond_tol(g, base=1e-8):
    cond = max(np.linalg.cond(L) ** 2 for L in g["L"])
    return max(base, 100 * cond * EPS)


@pytest.mark.parametrize("name", FILES)
def test_state_matches_reference(name):
    g = load(name)
    gp = build(g)
    for o in range(g["Y"].shape[1]):
        L = gp.gps[o].L_
        assert rel(L, g["L"][o]) < 1e-9
        K = g["L"][o] @ g["L"][o].T
        assert rel(K @ gp.gps[o].alpha_, g["Yn"][:, o]) < 1e-6
        assert abs(gp.gps[o].log_marginal_likelihood_value_ - g["lml_vals"][o, 1]) <= 1e-8 * abs(g["lml_vals"][o, 1]) + 1e-8
        assert np.allclose(gp.gps[o].kernel_.theta, g["theta"][o])
        assert np.allclose(gp.gps[o].kernel_.get_params()['k1__k2__length_scale'], np.exp(g["theta"][o][1:-1]))


@pytest.mark.parametrize("name", FILES)
def test_mean_and_gradient_vs_reference(name):
    """evaluate(std=False, gradient=True) batched — a valid batched call of the reference (gp.py:148-196)."""
    g = load(name)
    gp = build(g)
    out = gp.evaluate(g["Xs"], std=False, gradient=True)
    assert out["S"] is None and out["dS"] is None and out["hF"] is None
    tol = cond_tol(g)
    assert rel(out["F"], g["mean_F"]) < tol
    assert rel(out["dF"], g["mean_dF"]) < tol
    # a value-only call of this size runs in the fused small-batch kernel: same mean up to the summation order of
    # sum_j k_j alpha_j, whose cancellation grows with cond(K) like the tolerance against the reference does
    assert rel(gp.predict(g["Xs"]), out["F"]) < 1e-2 * tol


@pytest.mark.parametrize("name", FILES)
def test_std_vs_reference_and_oracle(name):
    g = load(name)
    gp = build(g)
    F, S = gp.predict(g["Xs"], std=True)
    tol = cond_tol(g)
    assert rel(F, g["std_F"]) < tol
    prior = np.exp(g["theta"][:, 0]) + np.exp(g["theta"][:, -1])
    var_gpu = (S / g["y_scale"]) ** 2
    var_ref = (g["std_S"] / g["y_scale"]) ** 2
    # sigma^2 is a cancellation against the prior variance: absolute error relative to c1+c2
    assert (np.abs(var_gpu - var_ref) / prior).max() < max(1e-10, 100 * max(np.linalg.cond(L) ** 2 for L in g["L"]) * EPS)
    # against the oracle's triangular form (same algorithm, same theta): tight
    orc = oracle_from(g)
    var_orc = (orc.evaluate(g["Xs"], std=True, var_form="chol")["S"] / g["y_scale"]) ** 2
    assert (np.abs(var_gpu - var_orc) / prior).max() < 1e-9


@pytest.mark.parametrize("name", FILES)
def test_std_gradient_rowstack_vs_reference(name):
    """Batched std+gradient == stack of the reference's per-row (C=1) results."""
    g = load(name)
    gp = build(g)
    out = gp.evaluate(g["Xs"], std=True, gradient=True)
    tol = cond_tol(g)
    assert rel(out["F"], g["row_F"]) < tol
    assert rel(out["dF"], g["row_dF"]) < tol
    prior = np.exp(g["theta"][:, 0]) + np.exp(g["theta"][:, -1])
    ok = (((g["row_S"] / g["y_scale"]) ** 2 / prior) > 1e-6).all(axis=1)
    cond = max(np.linalg.cond(L) ** 2 for L in g["L"])
    if cond < 1e9:
        assert rel(out["S"][ok], g["row_S"][ok]) < 1e-6
        assert rel(out["dS"][ok], g["row_dS"][ok]) < 1e-5
    # and tight against the oracle executing the same (triangular) algorithm
    orc = oracle_from(g)
    ref = orc.evaluate(g["Xs"], std=True, gradient=True, var_form="chol")
    assert rel(out["S"][ok], ref["S"][ok]) < 1e-7
    assert rel(out["dS"][ok], ref["dS"][ok]) < 1e-6


@pytest.mark.parametrize("name", FILES)
@pytest.mark.parametrize("kind", ["identity", "ucb", "ei", "pi"])
def test_acquisition_fused(name, kind, monkeypatch):
    import autooed_b200 as ab
    g = load(name)
    gp = build(g)
    cls = {"identity": ab.Identity, "ucb": ab.UpperConfidenceBound, "ei": ab.ExpectedImprovement,
           "pi": ab.ProbabilityOfImprovement}[kind]
    acq = cls(gp)
    acq.fit(g["X"], g["Y"])
    F, dF, hF = acq.evaluate(g["Xs"], gradient=True)
    assert hF is None
    # oracle transform applied to the GPU's own surrogate outputs: isolates the fused epilogue
    val = gp.evaluate(g["Xs"], std=True, gradient=True)
    val = dict(val, hF=None, hS=None)
    oF, odF, _ = go.acquisition(kind, val, True, False, n_sample=int(g["n_sample"]), y_min=g["y_min"])
    assert np.allclose(F, oF, rtol=1e-10, atol=1e-13)
    assert np.allclose(dF, odF, rtol=1e-9, atol=1e-12)
    # value-only batched call vs the reference (what NSGA-II issues every generation)
    Fb, dFb, _ = acq.evaluate(g["Xs"])
    assert dFb is None
    # ... on the chunked path the value does not depend on which derivatives were requested ...
    monkeypatch.setenv("AOED_SMALL_EVAL", "0")
    assert np.array_equal(acq.evaluate(g["Xs"])[0], F)
    monkeypatch.delenv("AOED_SMALL_EVAL")
    # ... and the fused small-batch kernel, which serves a call of this size, differs by the summation order only:
    # compared where sigma is not cancellation noise (the acquisitions divide by it)
    prior_all = np.exp(g["theta"][:, 0]) + np.exp(g["theta"][:, -1])
    solid = (((val["S"] / g["y_scale"]) ** 2 / prior_all) > 1e-6).all(axis=1)
    assert np.allclose(Fb[solid], F[solid], rtol=max(1e-9, 1e-2 * cond_tol(g)), atol=1e-12 * np.abs(F).max())
    cond = max(np.linalg.cond(L) ** 2 for L in g["L"])
    if cond < 1e9 or kind == "identity":
        prior = np.exp(g["theta"][:, 0]) + np.exp(g["theta"][:, -1])
        ok = (((g["row_S"] / g["y_scale"]) ** 2 / prior) > 1e-6).all(axis=1)
        # EI / PI tails (|z| ~ 20) magnify the reference's own sigma noise by z^2: 1e-4 there, 1e-5 elsewhere
        assert rel(Fb[ok], g[f"acq_{kind}_Fbatch"][ok]) < (1e-4 if kind in ("ei", "pi") else 1e-5)
    ex = cls(gp, exact=True)
    ex.fit(g["X"], g["Y"])
    eF, edF, _ = ex.evaluate(g["Xs"], gradient=True)
    xF, xdF, _ = go.acquisition(kind, val, True, False, n_sample=int(g["n_sample"]), y_min=g["y_min"], mode="exact")
    assert np.allclose(eF, xF, rtol=1e-10, atol=1e-13)
    assert np.allclose(edF, xdF, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("nu", [1, 3, 5, -1])
@pytest.mark.parametrize("shape", [(300, 20, 3, 700), (130, 7, 2, 129), (1, 3, 1, 5), (257, 32, 2, 300), (140, 40, 2, 300),
                                   (300, 64, 2, 150), (90, 33, 1, 40), (150, 70, 2, 60), (140, 100, 1, 50), (100, 128, 2, 30)])
def test_random_problem_vs_oracle(nu, shape, monkeypatch):
    """Multi-tile / multi-chunk shapes (ragged N, C; chunk of 256 rows forces the double-buffered path)."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    monkeypatch.setenv("AOED_CHUNK", "256")
    N, d, m, C = shape
    rng = np.random.default_rng(N + d + nu)
    X = rng.random((N, d))
    W = rng.standard_normal((d, m))
    Y = np.sin(X @ W) + 0.1 * (X ** 2) @ np.abs(W) + 0.05 * rng.standard_normal((N, m))
    thetas = np.stack([np.concatenate([[rng.uniform(-0.5, 0.5)], rng.uniform(-0.5, 1.0, d), [rng.uniform(-5, -3)]])
                       for _ in range(m)])
    Xs = rng.random((C, d))
    gp = GaussianProcess(SimpleProblem(d, m), nu=nu)
    gp.fit_fixed(X, Y, thetas)
    orc = go.GPOracle(d, m, nu=nu).fit(X, Y, thetas=thetas)
    out = gp.evaluate(Xs, std=True, gradient=True)
    ref = orc.evaluate(Xs, std=True, gradient=True, var_form="chol")
    cond = max(np.linalg.cond(st.L) ** 2 for st in orc.states)
    tol = max(1e-8, 100 * cond * EPS)
    assert rel(out["F"], ref["F"]) < tol
    assert rel(out["dF"], ref["dF"]) < tol
    prior = np.exp(thetas[:, 0]) + np.exp(thetas[:, -1])
    dv = np.abs((out["S"] / orc.norm.y_scale) ** 2 - (ref["S"] / orc.norm.y_scale) ** 2) / prior
    assert dv.max() < max(1e-10, tol * 1e-2)
    ok = ((ref["S"] / orc.norm.y_scale) ** 2 / prior > 1e-6).all(axis=1)
    if ok.any():
        assert rel(out["dS"][ok], ref["dS"][ok]) < max(1e-7, tol * 10)


def test_error_behaviour():
    from autooed_b200 import GaussianProcess, Identity
    from autooed_b200.problem import SimpleProblem
    gp = GaussianProcess(SimpleProblem(3, 2), nu=5)
    with pytest.raises(AssertionError):
        gp.evaluate(np.zeros((2, 3)))  # unfitted (base.py:94)
    with pytest.raises(AssertionError):
        Identity(gp).fit(np.zeros((2, 3)), np.zeros((2, 2)))  # acquisition/base.py:37
    rng = np.random.default_rng(0)
    X, Y = rng.random((10, 3)), rng.random((10, 2))
    gp.fit_fixed(X, Y, np.zeros((2, 5)))
    with pytest.raises(AssertionError):
        gp.evaluate(X, dtype='bogus')
    with pytest.raises(ValueError):
        gp.evaluate(X[0])  # 1-D input is illegal at the surrogate level
    out = gp.evaluate(np.empty((0, 3)), std=True, gradient=True)
    assert out["F"].shape == (0, 2) and out["dS"].shape == (0, 2, 3)


# ---- second derivatives (gp.py:209-242, ucb.py:44-51, ei.py:48-68, pi.py:47-62) -------------------------------------
@pytest.mark.parametrize("name", FILES)
def test_mean_hessian_batched_vs_reference(name):
    """evaluate(std=False, gradient=True, hessian=True) with C > 1 — a valid batched call of the reference."""
    g = load(name)
    gp = build(g)
    out = gp.evaluate(g["Xs"], std=False, gradient=True, hessian=True)
    assert out["hS"] is None and out["S"] is None
    tol = cond_tol(g)
    assert rel(out["F"], g["mean_F"]) < tol
    assert rel(out["dF"], g["mean_dF"]) < tol
    assert rel(out["hF"], g["mean_hF"]) < tol
    # hessian without gradient: same hF, no dF (the reference allows it for the mean)
    out2 = gp.evaluate(g["Xs"], hessian=True)
    assert out2["dF"] is None and np.array_equal(out2["hF"], out["hF"])


@pytest.mark.parametrize("name", FILES)
def test_full_hessian_rowstack(name):
    """std+gradient+hessian: stack of the reference's per-row (C=1) results, literal semantics."""
    g = load(name)
    gp = build(g)
    out = gp.evaluate(g["Xs"], std=True, gradient=True, hessian=True)
    tol = cond_tol(g)
    assert rel(out["hF"], g["row_hF"]) < tol
    orc = oracle_from(g)
    ref = orc.evaluate(g["Xs"], std=True, gradient=True, hessian=True, mode="literal", var_form="chol")
    prior = np.exp(g["theta"][:, 0]) + np.exp(g["theta"][:, -1])
    ok = (((g["row_S"] / g["y_scale"]) ** 2 / prior) > 1e-6).all(axis=1)
    assert ok.sum() >= 2
    assert rel(out["hF"], ref["hF"]) < tol  # alpha itself carries cond(K)*eps between two Cholesky implementations
    # same (triangular) algorithm as the oracle: tight, relative to the largest entry
    assert rel(out["hS"][ok], ref["hS"][ok]) < 1e-5
    cond = max(np.linalg.cond(L) ** 2 for L in g["L"])
    if cond < 1e9:  # where the reference's own explicit-inverse variance keeps its digits
        assert rel(out["hS"][ok], g["row_hS"][ok]) < 1e-4
    # std + hessian without gradient: the reference raises UnboundLocalError; here it is the same hS
    out2 = gp.evaluate(g["Xs"], std=True, hessian=True)
    assert out2["dS"] is None and np.array_equal(out2["hS"], out["hS"])


@pytest.mark.parametrize("name", FILES)
@pytest.mark.parametrize("kind", ["identity", "ucb", "ei", "pi"])
@pytest.mark.parametrize("exact", [False, True])
def test_acquisition_hessian_fused(name, kind, exact):
    """Fused acquisition Hessian == the oracle's transform applied to the GPU's own surrogate outputs."""
    import autooed_b200 as ab
    g = load(name)
    gp = build(g)
    cls = {"identity": ab.Identity, "ucb": ab.UpperConfidenceBound, "ei": ab.ExpectedImprovement,
           "pi": ab.ProbabilityOfImprovement}[kind]
    acq = cls(gp, exact=exact)
    acq.fit(g["X"], g["Y"])
    F, dF, hF = acq.evaluate(g["Xs"], gradient=True, hessian=True)
    val = gp._evaluate(gp.normalization.do(x=g["Xs"]), True, True, True, exact=exact)
    oF, odF, ohF = go.acquisition(kind, val, True, True, n_sample=int(g["n_sample"]), y_min=g["y_min"],
                                  mode="exact" if exact else "literal")
    fin = np.isfinite(ohF)
    assert (np.isfinite(hF) == fin).all()
    assert np.allclose(F, oF, rtol=1e-10, atol=1e-13)
    assert np.allclose(dF, odF, rtol=1e-9, atol=1e-12)
    assert np.allclose(hF[fin], ohF[fin], rtol=1e-8, atol=1e-10 * max(1.0, np.abs(ohF[fin]).max()))
    if not exact and kind in ("identity", "ucb"):
        cond = max(np.linalg.cond(L) ** 2 for L in g["L"])
        prior = np.exp(g["theta"][:, 0]) + np.exp(g["theta"][:, -1])
        ok = (((g["row_S"] / g["y_scale"]) ** 2 / prior) > 1e-6).all(axis=1)
        if kind == "identity":
            assert rel(hF, g["acq_identity_hF"]) < cond_tol(g)
        elif cond < 1e9:
            assert rel(hF[ok], g["acq_ucb_hF"][ok]) < 1e-4


@pytest.mark.parametrize("nu", [1, 3, 5, -1])
@pytest.mark.parametrize("shape", [(300, 20, 3, 37), (130, 7, 2, 700), (257, 32, 2, 9), (120, 40, 2, 11), (150, 64, 1, 5), (90, 100, 1, 3)])
def test_exact_hessian_vs_oracle_and_fd(nu, shape, monkeypatch):
    """`exact` mode: true second derivatives in raw-Y units; multi-chunk (rows*d > chunk) shapes."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    monkeypatch.setenv("AOED_CHUNK", "512")
    N, d, m, C = shape
    rng = np.random.default_rng(7 * N + d + nu)
    X = rng.random((N, d))
    W = rng.standard_normal((d, m))
    Y = np.sin(X @ W) + 0.1 * (X ** 2) @ np.abs(W) + 0.05 * rng.standard_normal((N, m))
    thetas = np.stack([np.concatenate([[rng.uniform(-0.5, 0.5)], rng.uniform(-0.5, 1.0, d), [rng.uniform(-5, -3)]])
                       for _ in range(m)])
    Xs = rng.random((C, d))
    gp = GaussianProcess(SimpleProblem(d, m), nu=nu)
    gp.fit_fixed(X, Y, thetas)
    orc = go.GPOracle(d, m, nu=nu).fit(X, Y, thetas=thetas)
    for mode in ("literal", "exact"):
        out = gp._evaluate(Xs, True, True, True, exact=(mode == "exact"))
        ref = orc.evaluate(Xs, std=True, gradient=True, hessian=True, mode=mode, var_form="chol", normalized=True)
        cond = max(np.linalg.cond(st.L) ** 2 for st in orc.states)
        tol = max(1e-8, 100 * cond * EPS)
        assert rel(out["hF"], ref["hF"]) < tol
        prior = np.exp(thetas[:, 0]) + np.exp(thetas[:, -1])
        ok = ((ref["S"] / orc.norm.y_scale) ** 2 / prior > 1e-6).all(axis=1)
        assert rel(out["hS"][ok], ref["hS"][ok]) < max(1e-6, tol * 100)
    if nu != 1:  # exact mode is the true Hessian: central differences of the GPU gradient (nu=1/2 is not C^2 at data)
        h = 1e-5
        out = gp._evaluate(Xs[:4], True, True, True, exact=True)
        for k in range(min(d, 3)):
            e = np.zeros(d); e[k] = h
            op = gp._evaluate(Xs[:4] + e, True, True, False)
            om = gp._evaluate(Xs[:4] - e, True, True, False)
            assert np.allclose((op["dF"] - om["dF"]) / (2 * h), out["hF"][:, :, :, k], rtol=1e-4, atol=1e-5 * np.abs(out["hF"]).max())
            assert np.allclose((op["dS"] - om["dS"]) / (2 * h), out["hS"][:, :, :, k], rtol=1e-3, atol=1e-4 * np.abs(out["hS"]).max())


@pytest.mark.parametrize("N", [2, 15, 16, 17, 127, 128, 129, 255, 256, 257, 511, 513, 1025])
def test_tile_boundaries_vs_oracle(N):
    """Training-set sizes around every tiling boundary of the triangular products (k-step 16, row tile 128, sum-of-squares
    slices of 32 rows) and candidate counts around the 128-column tile, std + gradient + literal Hessian."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    d, m = 5, 2
    rng = np.random.default_rng(1000 + N)
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, m))) + 0.05 * rng.standard_normal((N, m))
    thetas = np.stack([np.concatenate([[0.2], rng.uniform(-0.8, 0.3, d), [-4.0]]) for _ in range(m)])
    gp = GaussianProcess(SimpleProblem(d, m), nu=3)
    gp.fit_fixed(X, Y, thetas)
    orc = go.GPOracle(d, m, nu=3).fit(X, Y, thetas=thetas)
    cond = max(np.linalg.cond(st.L) ** 2 for st in orc.states)
    tol = max(1e-8, 100 * cond * EPS)
    prior = np.exp(thetas[:, 0]) + np.exp(thetas[:, -1])
    for C in (1, 127, 129):
        Xs = rng.random((C, d))
        out = gp.evaluate(Xs, std=True, gradient=True, hessian=C == 1)
        ref = orc.evaluate(Xs, std=True, gradient=True, hessian=C == 1, var_form="chol")
        assert rel(out["F"], ref["F"]) < tol and rel(out["dF"], ref["dF"]) < tol
        dv = np.abs((out["S"] / orc.norm.y_scale) ** 2 - (ref["S"] / orc.norm.y_scale) ** 2) / prior
        assert dv.max() < max(1e-10, tol * 1e-2)
        ok = ((ref["S"] / orc.norm.y_scale) ** 2 / prior > 1e-6).all(axis=1)
        if ok.any():
            assert rel(out["dS"][ok], ref["dS"][ok]) < max(1e-7, tol * 10)
            if C == 1:
                assert rel(out["hF"], ref["hF"]) < tol
                assert rel(out["hS"][ok], ref["hS"][ok]) < max(1e-6, tol * 100)


@pytest.mark.parametrize("name", [f for f in FILES if f.startswith(("zdt1", "dtlz2"))])
def test_variance_closer_to_truth_than_reference(name):
    """SURVEY.md B-6 / Appendix E contract (ii): on noise-free fits cond(K) reaches 1e10 .. 1e13 and the reference's
    explicit-inverse variance is cancellation noise; the CUDA path (triangular form, FP64 DMMA) must be at least as close to
    an 80-bit ground truth as the reference is, and within 1e-11 of the prior variance in absolute terms."""
    g = load(name)
    gp = build(g)
    _, S = gp.predict(g["Xs"], std=True)
    prior = np.exp(g["theta"][:, 0]) + np.exp(g["theta"][:, -1])
    for o in range(g["Y"].shape[1]):
        truth = go.variance_truth_longdouble(g["theta"][o], g["Xn"], (g["Xs"] - g["xl"]) / (g["xu"] - g["xl"]), int(g["nu"]))
        truth = np.maximum(truth, 0.0)
        var_gpu = (S[:, o] / g["y_scale"][o]) ** 2
        var_ref = (g["std_S"][:, o] / g["y_scale"][o]) ** 2
        err_gpu, err_ref = np.abs(var_gpu - truth), np.abs(var_ref - truth)
        assert err_gpu.max() <= 1e-11 * prior[o], (o, err_gpu.max(), prior[o])
        assert err_gpu.max() <= max(err_ref.max(), 1e-13 * prior[o]), (o, err_gpu.max(), err_ref.max())


@pytest.mark.parametrize("nu", [1, 3, 5, -1])
@pytest.mark.parametrize("N,d,rows", [(40, 6, 1), (120, 6, 1000), (129, 20, 37), (256, 32, 530), (7, 3, 4096), (200, 48, 77),
                                      (256, 64, 300), (100, 100, 64), (256, 128, 40)])
def test_small_batch_kernel_equals_chunked_path(nu, N, d, rows, monkeypatch):
    """Value-only calls with n_train <= 256 and <= 4096 rows run in one fused kernel (csrc/small_eval.cuh);
    AOED_SMALL_EVAL=0 sends the same call down the chunked xcov / triangular-product / epilogue path.  Same model, same
    candidates: the two agree to rounding (different summation orders), for the surrogate outputs and every acquisition."""
    import autooed_b200 as ab
    from autooed_b200.problem import SimpleProblem
    rng = np.random.default_rng(N + rows)
    m = 2
    X, Y = rng.random((N, d)), rng.standard_normal((N, m))
    gp = ab.GaussianProcess(SimpleProblem(d, m), nu=nu)
    theta = np.tile(np.concatenate([[0.3], np.log(rng.uniform(0.5, 2.0, d)), [-3.0]]), (m, 1))
    gp.fit_fixed(X, Y, theta)
    Xs = rng.random((rows, d))
    Xs[0] = X[0]  # zero distance / near-zero variance
    prior = np.exp(theta[0, 0]) + np.exp(theta[0, -1])
    acqs = [ab.Identity(gp), ab.UpperConfidenceBound(gp), ab.ExpectedImprovement(gp), ab.ProbabilityOfImprovement(gp)]
    for a in acqs:
        a.fit(X, Y)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("AOED_SMALL_EVAL", mode)
        out = gp.evaluate(Xs, std=True)
        res[mode] = [out["F"], out["S"], gp.evaluate(Xs)["F"]] + [a.evaluate(Xs)[0] for a in acqs]
    y_scale = Y.std(axis=0)
    scale = np.abs(res["0"][0]).max()
    # mean = sum_j k_j alpha_j: the two paths add the N terms in different orders, each term is up to prior * |alpha|
    amax = max(np.abs(v.alpha_).max() for v in gp.gps)
    mean_tol = max(1e-12 * scale, 4 * N * np.finfo(float).eps * amax * prior * y_scale.max())
    assert np.abs(res["1"][0] - res["0"][0]).max() <= mean_tol
    assert np.array_equal(res["1"][2], res["1"][0])                  # std=False call gives the same mean
    # sigma = sqrt(prior - |w|^2): compare variances, the quantity the two paths sum differently
    v1, v0 = (res["1"][1] / y_scale) ** 2, (res["0"][1] / y_scale) ** 2
    assert np.abs(v1 - v0).max() <= 1e-12 * prior
    ok = v0.min(axis=1) > 1e-6 * prior  # acquisitions divide by sigma: compare where sigma is not cancellation noise
    for k in range(3, 7):
        assert np.allclose(res["1"][k][ok], res["0"][k][ok], rtol=1e-8, atol=1e-10 * scale + 10 * mean_tol)


def test_fused_gradient_moments_equal_separate_kernel(tmp_path):
    """AOED_GRADSUM=fused (variance-gradient moments formed in the epilogue of the UPPER triangular product) against the
    default gradsum kernel: same dS / dA up to summation order.  The switch is latched per process: child process."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = f"""
import sys, numpy as np
sys.path[:0] = [{here!r}, {os.path.dirname(here)!r}]
from autooed_b200 import GaussianProcess, UpperConfidenceBound
from autooed_b200.problem import SimpleProblem
rng = np.random.default_rng(0)
N, d, m = 300, 7, 2
X = rng.random((N, d)); Y = np.sin(X @ rng.standard_normal((d, m))) + 0.05 * rng.standard_normal((N, m))
th = np.tile(np.concatenate([[0.0], np.full(d, -0.3), [-4.0]]), (m, 1))
gp = GaussianProcess(SimpleProblem(d, m), nu=5); gp.fit_fixed(X, Y, th)
acq = UpperConfidenceBound(gp); acq.fit(X, Y)
Xs = rng.random((700, d))
out = gp.evaluate(Xs, std=True, gradient=True)
A, dA, _ = acq.evaluate(Xs, gradient=True)
np.savez(sys.argv[1], S=out['S'], dS=out['dS'], dF=out['dF'], A=A, dA=dA)
"""
    res = {}
    for mode in ("separate", "fused"):
        path = str(tmp_path / f"{mode}.npz")
        env = dict(os.environ, AOED_GRADSUM=mode)
        r = subprocess.run([sys.executable, "-c", code, path], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        res[mode] = dict(np.load(path))
    for k in ("S", "dF", "A"):
        assert np.array_equal(res["fused"][k], res["separate"][k]), k
    for k in ("dS", "dA"):
        assert np.allclose(res["fused"][k], res["separate"][k], rtol=1e-9, atol=1e-12), k
    assert np.abs(res["fused"]["dS"]).max() > 0
