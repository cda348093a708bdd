"""C-ABI error behaviour on a real device, and the reference's preset algorithms running end to end on the B200 path
(TSEMO = gp + ts + nsga2 + hvi, the CLI default; USeMO-EI = gp + ei + nsga2 + uncertainty)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cabi_error_codes():
    from autooed_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.aoed_gp_create(ctypes.byref(h), 0, 129, 2, 5) == _lib.ERR_UNSUPPORTED  # n_var > 128
    assert b"n_var" in lib.aoed_last_error()
    assert lib.aoed_gp_create(ctypes.byref(h), 0, 4, 2, 7) == _lib.ERR_INVALID       # nu not in {1, 3, 5, -1}
    assert lib.aoed_gp_create(ctypes.byref(h), 99, 4, 2, 5) == _lib.ERR_NO_DEVICE
    assert lib.aoed_gp_create(ctypes.byref(h), 0, 4, 2, 5) == _lib.OK
    X = np.zeros((3, 4))
    out = np.zeros((3, 2))
    rc = lib.aoed_gp_eval(h, 3, _lib.ptr(X), 0, 0, None, _lib.ptr(out), *([None] * 8))
    assert rc == _lib.ERR_INVALID and b"not fitted" in lib.aoed_last_error()  # evaluate before fit (base.py:94)
    th = np.zeros((2, 6))
    assert lib.aoed_gp_set_theta(h, _lib.ptr(th)) == _lib.ERR_INVALID  # set_train first
    Xn, Yn = np.random.default_rng(0).random((10, 4)), np.random.default_rng(1).random((10, 2))
    assert lib.aoed_gp_set_train(h, 10, _lib.ptr(Xn), _lib.ptr(Yn), None, None) == _lib.OK
    bad = np.array([[0, 0, 0, 0, 0, 0.0]] * 2)
    lml = np.zeros(1)
    assert lib.aoed_gp_lml_batch(h, 1, _lib.ptr(np.array([5], dtype=np.int32)), _lib.ptr(bad), _lib.ptr(lml), None) == _lib.ERR_INVALID
    assert lib.aoed_gp_set_theta(h, _lib.ptr(th)) == _lib.OK
    rc = lib.aoed_gp_eval(h, 3, _lib.ptr(X), 1, 2, None, *([None] * 6), _lib.ptr(out), None, None)
    assert rc == _lib.ERR_INVALID  # UCB without its parameter
    rc = lib.aoed_gp_eval(h, 3, _lib.ptr(X), 1, 9, None, *([None] * 6), _lib.ptr(out), None, None)
    assert rc == _lib.ERR_INVALID  # unknown acquisition kind
    assert lib.aoed_gp_eval(h, -1, _lib.ptr(X), 0, 0, None, _lib.ptr(out), *([None] * 8)) == _lib.ERR_INVALID
    assert lib.aoed_gp_eval(h, 3, _lib.ptr(X), 0, 0, None, _lib.ptr(out), *([None] * 8)) == _lib.OK
    assert lib.aoed_gp_destroy(h) == _lib.OK


def test_non_pd_final_fit_raises_linalgerror():
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    from oracle import gp_oracle as go
    # duplicated rows, smooth kernel, and a signal variance (e^25) whose rounding error swamps the 1e-10 jitter: LAPACK
    # fails on the CPU (sklearn raises LinAlgError, _gpr.py:352-361) and so must the device factorisation
    X = np.repeat(np.random.default_rng(0).random((5, 3)), 8, axis=0)
    Y = np.random.default_rng(1).random((40, 1))
    theta = np.array([[25.0, 3.4, 3.4, 3.4, -6.0]])
    with pytest.raises(np.linalg.LinAlgError):
        go.GPState(theta[0], X, (Y[:, 0] - Y.mean()) / Y.std(), -1)
    gp = GaussianProcess(SimpleProblem(3, 1), nu=-1)
    with pytest.raises(np.linalg.LinAlgError):
        gp.fit_fixed(X, Y, theta)


def _zdt1(X):
    f1 = X[:, 0]
    g = 1 + 9.0 / (X.shape[1] - 1) * X[:, 1:].sum(1)
    return np.stack([f1, g * (1 - np.sqrt(f1 / g))], 1)


@pytest.mark.parametrize("spec", [("ts", "hvi"), ("ei", "uncertainty")])
def test_preset_algorithms_end_to_end(spec):
    import autooed_b200 as ab
    from autooed_b200.mobo import MOBO
    from autooed_b200.problem import SimpleProblem
    from autooed_b200.utils.sampling import lhs
    from oracle import moo_oracle as mo
    acq_name, sel_name = spec
    np.random.seed(0)
    prob = SimpleProblem(6, 2)
    X = lhs(6, 24)
    Y = _zdt1(X)
    sur = ab.init_surrogate_model('gp', {'nu': 5}, prob)
    acq = ab.init_acquisition(acq_name, {}, sur)
    solver = ab.init_solver('nsga2', {'n_gen': 40, 'pop_size': 100}, prob)
    sel = ab.init_selection(sel_name, {}, sur)
    opt = MOBO(prob, sur, acq, solver, sel)
    hv0 = mo.hypervolume(Y, [1.1, 7.0])
    for _ in range(3):
        Xn = opt.optimize(X, Y, None, 6)
        assert Xn.shape == (6, 6) and Xn.min() >= 0.0 and Xn.max() <= 1.0
        X, Y = np.vstack([X, Xn]), np.vstack([Y, _zdt1(Xn)])
    assert mo.hypervolume(Y, [1.1, 7.0]) > hv0
