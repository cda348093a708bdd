"""TEST INFRASTRUCTURE ONLY — imports the reference's own GP / acquisition modules from
/root/reference (read-only mount; exists only in the authoring container, never on the GPU box).

Used by the `oracle/make_golden*.py` scripts to generate the committed fixtures under `tests/golden/` (replayed by
`tests/test_oracle_golden.py` and the GPU tests).  Nothing in the
product package may import this module.

Why a shim: `autooed/mobo/__init__.py:1` imports `mobo.py` -> `factory.py` -> `solver/nsga2.py:6`,
which needs pymoo 0.4.1 (not installable offline).  We pre-register empty package stubs whose
`__path__` points at the real directories, so the leaf modules
(`surrogate_model/gp.py`, `acquisition/{identity,ei,pi,ucb}.py`) import unmodified.
After `fit`, sklearn >= 0.24 no longer defines `_K_inv`, which `gp.py:154` reads; we set it to None.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("AOED_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "autooed", "mobo", "surrogate_model"))


def _stub(name, path):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod
    return mod


def load():
    """Return a namespace with the reference classes (GaussianProcess, acquisitions, ...)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    base = os.path.join(REF_ROOT, "autooed")
    _stub("autooed", base)
    _stub("autooed.mobo", os.path.join(base, "mobo"))
    _stub("autooed.mobo.surrogate_model", os.path.join(base, "mobo", "surrogate_model"))
    _stub("autooed.mobo.acquisition", os.path.join(base, "mobo", "acquisition"))
    _stub("autooed.utils", os.path.join(base, "utils"))
    ns = types.SimpleNamespace()
    ns.gp = importlib.import_module("autooed.mobo.surrogate_model.gp")
    ns.GaussianProcess = ns.gp.GaussianProcess
    ns.Matern = ns.gp.Matern
    ns.Identity = importlib.import_module("autooed.mobo.acquisition.identity").Identity
    ns.ExpectedImprovement = importlib.import_module("autooed.mobo.acquisition.ei").ExpectedImprovement
    ns.ProbabilityOfImprovement = importlib.import_module("autooed.mobo.acquisition.pi").ProbabilityOfImprovement
    ns.UpperConfidenceBound = importlib.import_module("autooed.mobo.acquisition.ucb").UpperConfidenceBound
    ns.normalization = importlib.import_module("autooed.utils.normalization")
    ns.safe_divide = importlib.import_module("autooed.utils.operand").safe_divide
    return ns


class FakeTransformation:
    """Continuous transformation (`problem/transformation.py:46-52`) without the Problem machinery."""

    def do(self, X):
        import numpy as np
        return np.array(X, dtype=object).astype(float)

    def undo(self, X):
        import numpy as np
        return np.array(X, dtype=float)


class FakeProblem:
    """Only what `SurrogateModel.__init__` reads (`surrogate_model/base.py:24-28`)."""

    def __init__(self, n_var, n_obj, xl=None, xu=None):
        import numpy as np
        self.n_var, self.n_obj, self.n_constr = n_var, n_obj, 0
        self.xl = np.zeros(n_var) if xl is None else np.asarray(xl, float)
        self.xu = np.ones(n_var) if xu is None else np.asarray(xu, float)
        self.transformation = FakeTransformation()
        self.obj_type = ["min"] * n_obj


def fit_reference_gp(ns, problem, X, Y, nu):
    """Fit the reference GaussianProcess and apply the `_K_inv` attribute shim."""
    model = ns.GaussianProcess(problem, nu=nu)
    model.fit(X, Y)
    for g in model.gps:
        g._K_inv = None
    return model
