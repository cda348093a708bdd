"""CPU-side checks of the boundary: the shared library builds, loads, exports every symbol that
include/aoed_b200.h declares, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from autooed_b200 import _lib
    return _lib


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "aoed_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aoed_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/aoed_b200.h but not exported"
    assert sorted(lib.SIGNATURES) == names, "ctypes signature table out of sync with the header"
    assert lib.load().aoed_abi_version() == 1


def test_no_cpu_fallback(lib):
    if lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    from autooed_b200.utils import pareto
    gp = GaussianProcess(SimpleProblem(3, 2))
    with pytest.raises(lib.AoedError):
        gp.fit(np.random.rand(8, 3), np.random.rand(8, 2))
    with pytest.raises(lib.AoedError):
        pareto.check_pareto(np.random.rand(4, 2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "autooed_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no CPU fallback", ""), f"{f} mentions the oracle"


def test_host_side_normalization_matches_oracle():
    from autooed_b200.utils.normalization import StandardNormalization
    from oracle import gp_oracle as go
    rng = np.random.default_rng(0)
    xl, xu = np.array([-1.0, 2.0]), np.array([3.0, 5.0])
    X = xl + rng.random((20, 2)) * (xu - xl) * 1.2
    Y = np.hstack([rng.random((20, 1)), np.full((20, 1), 3.0)])  # constant column -> scale 1
    n = StandardNormalization(np.array([xl, xu]))
    n.fit(X, Y)
    o = go.Normalization(xl, xu)
    o.fit(Y)
    assert np.array_equal(n.do(x=X), o.do_x(X))
    assert np.allclose(n.do(y=Y), o.do_y(Y), rtol=0, atol=0)
    assert n.y_scaler.scale_[1] == 1.0
    assert np.allclose(n.undo(y=n.do(y=Y)), Y)
    assert np.allclose(n.rescale(y=np.ones((3, 2))), n.y_scaler.scale_ * np.ones((3, 2)))


def test_lockstep_driver_batches_all_problems():
    """The L-BFGS-B lock-step driver (host logic) on a cheap analytic objective: every batch carries all live problems."""
    from autooed_b200.surrogate_model.gp import _LockstepLML
    import threading
    from scipy.optimize import minimize
    sizes = []

    def evaluate_batch(ids, thetas):
        sizes.append(len(ids))
        centers = np.array([[b, -b] for b in ids], dtype=float)
        lml = -((thetas - centers) ** 2).sum(1)
        return lml, -2 * (thetas - centers)

    n = 5
    pool = _LockstepLML(evaluate_batch, n)
    res = [None] * n

    def worker(b):
        try:
            res[b] = minimize(lambda th: pool.objective(b, th), np.zeros(2), method="L-BFGS-B", jac=True).x
        finally:
            pool.finish(b)

    ts = [threading.Thread(target=worker, args=(b,)) for b in range(n)]
    [t.start() for t in ts]
    pool.serve()
    [t.join() for t in ts]
    for b in range(n):
        assert np.allclose(res[b], [b, -b], atol=1e-5)
    assert sizes[0] == n and max(sizes) == n


def test_argument_checks_need_no_device(lib):
    """Invalid arguments are rejected before any CUDA call (the same codes on a box without a GPU)."""
    import ctypes
    cdll = lib.load()
    F = np.zeros((4, 2))
    rank, crowd, sel = np.zeros(4, np.int32), np.zeros(4), np.zeros(2, np.int32)
    args = (lib.ptr(F), 2, lib.ptr(rank), lib.ptr(crowd), lib.ptr(sel))
    assert cdll.aoed_nsga2_survival(0, 0, 2, *args) == lib.ERR_INVALID                       # no individuals
    assert cdll.aoed_nsga2_survival(0, 4, 2, lib.ptr(F), 5, *args[2:]) == lib.ERR_INVALID    # keep > n
    assert cdll.aoed_nsga2_survival(0, 4, 2, None, 2, *args[2:]) == lib.ERR_INVALID          # null objectives
    assert cdll.aoed_nsga2_survival(0, 5000, 2, *args) == lib.ERR_UNSUPPORTED                # above the 4096-row limit
    assert cdll.aoed_nsga2_survival(0, 4, 9, *args) == lib.ERR_UNSUPPORTED                   # more than 8 objectives
    assert b"4096" in cdll.aoed_last_error()
    n_out, n_eval = ctypes.c_int64(), ctypes.c_int64()
    x = np.zeros((4, 3))
    rc = cdll.aoed_nsga2_run(None, 0, 3, 2, 1, None, 4, lib.ptr(x), lib.ptr(x), lib.ptr(x), 8, 5, 1, lib.ptr(x), lib.ptr(F),
                             ctypes.byref(n_out), ctypes.byref(n_eval))
    assert rc == lib.ERR_INVALID  # null model handle
