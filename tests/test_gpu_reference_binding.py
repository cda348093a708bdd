"""The reference's OWN calling code bound to the B200 surrogate — the drop-in claim of INTEGRATION.md §1, executed.

`__graft_entry__.build()` stages the unmodified reference modules `autooed/mobo/acquisition/{base,identity,ei,pi,ucb,ts,
penalize}.py` and the `autooed/utils` they import under the git-ignored `baseline/_ref/` (they travel to the GPU box with
the snapshot; the test is skipped where they are absent).  Here they are imported as they are, with the one change
a maintainer makes in `autooed/mobo/surrogate_model/__init__.py` / `gp.py`: `GaussianProcess` names
`autooed_b200.GaussianProcess`.  Their `fit` / `evaluate` then run
    surrogate_model.evaluate(X, dtype='continuous', std=True, gradient=..., hessian=...)          (ei.py:24, ucb.py:23, pi.py:24)
    surrogate_model.gps[i].fit(X, y), .kernel_.theta                                                 (ts.py:31-38)
    surrogate_model.gps[i].kernel_.get_params()['k1__k2__length_scale'], .predict(X_busy, std=True)  (penalize.py:141,165)
against the GPU model, and the results must equal the fused device acquisitions of this package."""
import importlib
import os
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def _ref():
    """Imports the staged reference acquisition modules on top of stub packages (the real `autooed/mobo/__init__.py` pulls
    pymoo through the solver factory, which is irrelevant here)."""
    base = os.path.join(REF, "autooed")
    if not os.path.isdir(os.path.join(base, "mobo", "acquisition")):
        pytest.skip("baseline/_ref is not staged (run __graft_entry__.build() where /root/reference is mounted)")
    import autooed_b200 as ab

    def stub(name, path=None):
        mod = sys.modules.get(name)
        if mod is None or not getattr(mod, "_aoed_stub", False):
            mod = types.ModuleType(name)
            mod._aoed_stub = True
            if path is not None:
                mod.__path__ = [path]
            sys.modules[name] = mod
        return mod

    stub("autooed", base)
    stub("autooed.mobo", os.path.join(base, "mobo"))
    stub("autooed.mobo.acquisition", os.path.join(base, "mobo", "acquisition"))
    stub("autooed.utils", os.path.join(base, "utils"))
    # the binding: the surrogate the reference modules import IS the B200 one
    sm = stub("autooed.mobo.surrogate_model", None)
    sm.GaussianProcess = ab.GaussianProcess
    smgp = stub("autooed.mobo.surrogate_model.gp", None)
    smgp.GaussianProcess = ab.GaussianProcess
    ns = types.SimpleNamespace()
    for mod, names in (("identity", ["Identity"]), ("ei", ["ExpectedImprovement"]), ("pi", ["ProbabilityOfImprovement"]),
                       ("ucb", ["UpperConfidenceBound"]), ("ts", ["ThompsonSampling"]),
                       ("penalize", ["LocalPenalization", "LocalLipschitzPenalization", "HardLocalPenalization"])):
        m = importlib.import_module(f"autooed.mobo.acquisition.{mod}")
        assert os.path.realpath(m.__file__).startswith(os.path.realpath(REF))
        for n in names:
            setattr(ns, n, getattr(m, n))
    return ns


def _model(nu=5, N=60, d=5, m=2, seed=0):
    import autooed_b200 as ab
    from autooed_b200.problem import SimpleProblem
    rng = np.random.default_rng(seed)
    X = rng.random((N, d))
    Y = np.stack([np.sin(3 * X[:, 0]) + X[:, 1] ** 2 + 0.05 * rng.standard_normal(N),
                  np.cos(2 * X[:, 2]) * X[:, 3] + X[:, 4] + 0.05 * rng.standard_normal(N)], 1)[:, :m]
    gp = ab.GaussianProcess(SimpleProblem(d, m), nu=nu)
    gp.fit(X, Y)
    return gp, X, Y, rng.random((40, d))


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("nu", [1, 5])
@pytest.mark.parametrize("name", ["Identity", "UpperConfidenceBound", "ExpectedImprovement", "ProbabilityOfImprovement"])
def test_reference_acquisition_classes_drive_the_b200_surrogate(name, nu):
    import autooed_b200 as ab
    ref = _ref()
    gp, X, Y, Xs = _model(nu=nu)
    theirs, ours = getattr(ref, name)(gp), getattr(ab, name)(gp)
    theirs.fit(X, Y)
    ours.fit(X, Y)
    # value: the reference evaluates whole batches
    F_ref, _, _ = theirs.evaluate(Xs, dtype='continuous')
    F, dF, hF = ours.evaluate(Xs, dtype='continuous', gradient=True, hessian=True)
    assert rel(F, F_ref) < 1e-10
    # gradient / Hessian: the reference's std-derivatives are only defined row by row
    for i in range(8):
        Fi, dFi, hFi = theirs.evaluate(Xs[i:i + 1], dtype='continuous', gradient=True, hessian=True)
        assert rel(F[i:i + 1], Fi) < 1e-10
        assert rel(dF[i:i + 1], dFi) < 1e-9
        assert rel(hF[i:i + 1], hFi) < 1e-8


def test_reference_thompson_sampling_refits_through_gps_and_draws_the_same_sample():
    import autooed_b200 as ab
    ref = _ref()
    gp, X, Y, Xs = _model(nu=5)
    theta_before = np.stack([g.kernel_.theta for g in gp.gps])
    theirs, ours = ref.ThompsonSampling(gp), ab.ThompsonSampling(gp)  # ts.py:21 isinstance(GaussianProcess) holds
    np.random.seed(4)
    theirs.fit(X, Y)  # ts.py:31-38: gps[i].fit(X, y), kernel_.theta
    np.random.seed(4)
    ours.fit(X, Y)
    assert np.array_equal(np.stack([g.kernel_.theta for g in gp.gps]), theta_before)  # the refit of the same data is the same fit
    for Wr, Wo, br, bo in zip(theirs.Ws, ours.Ws, theirs.bs, ours.bs):
        assert np.array_equal(Wr, Wo) and np.array_equal(br, bo)
    F_ref, dF_ref, hF_ref = theirs.evaluate(Xs, dtype='continuous', gradient=True, hessian=True)
    F, dF, hF = ours.evaluate(Xs, dtype='continuous', gradient=True, hessian=True)
    assert rel(F, F_ref) < 1e-10 and rel(dF, dF_ref) < 1e-10 and rel(hF, hF_ref) < 1e-10


@pytest.mark.parametrize("kind", ["LocalPenalization", "LocalLipschitzPenalization", "HardLocalPenalization"])
def test_reference_penalisers_on_the_b200_surrogate(kind):
    import autooed_b200 as ab
    from autooed_b200.acquisition import penalize as pz
    ref = _ref()
    gp, X, Y, Xs = _model(nu=3)
    X_busy = np.random.default_rng(1).random((3, X.shape[1]))
    theirs = getattr(ref, kind)(ref.Identity(gp))
    ours = getattr(pz, kind)(ab.Identity(gp))
    np.random.seed(2)
    theirs.fit(X, Y, X_busy)  # get_params()['k1__k2__length_scale'], predict(X_busy, std=True), evaluate(gradient=True)
    np.random.seed(2)
    ours.fit(X, Y, X_busy)
    F_ref, _, _ = theirs.evaluate(Xs, dtype='continuous')
    F, _, _ = ours.evaluate(Xs, dtype='continuous')
    assert rel(F, F_ref) < 1e-4  # the Lipschitz estimate is an L-BFGS-B polish with numerical gradients (2e-5 stopping tolerance)
