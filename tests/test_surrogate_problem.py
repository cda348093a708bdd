"""Solver-facing adapter (`autooed/mobo/surrogate_problem.py:54-193` semantics): flag decoding, 1-D inputs,
constraints — host logic on CPU with a stub acquisition, and the real GPU acquisition behind it."""
import numpy as np
import pytest

from autooed_b200.problem import SimpleProblem
from autooed_b200.surrogate_problem import SurrogateProblem


class StubAcq:
    fitted = True

    def __init__(self, m, d):
        self.m, self.d, self.calls = m, d, []

    def evaluate(self, X, dtype='raw', gradient=False, hessian=False):
        self.calls.append((dtype, gradient, hessian))
        n = len(X)
        F = np.tile(X.sum(1, keepdims=True), (1, self.m))
        dF = np.ones((n, self.m, self.d)) if gradient else None
        hF = np.zeros((n, self.m, self.d, self.d)) if hessian else None
        return F, dF, hF


class Constrained(SimpleProblem):
    def evaluate_constraint(self, x):
        return np.array([x[0] - 0.5, -1.0])


def test_flag_decoding_and_shapes():
    acq = StubAcq(2, 3)
    sp = SurrogateProblem(SimpleProblem(3, 2), acq)
    X = np.random.default_rng(0).random((5, 3))
    F = sp.evaluate(X)
    assert F.shape == (5, 2) and acq.calls[-1] == ('continuous', False, False)
    out = sp.evaluate(X, return_values_of=['F', 'dF', 'hF', 'CV', 'feasible'], return_as_dictionary=True)
    assert acq.calls[-1] == ('continuous', True, True)
    assert out['dF'].shape == (5, 2, 3) and out['hF'].shape == (5, 2, 3, 3)
    assert np.array_equal(out['CV'], np.zeros((5, 1))) and out['feasible'].all()  # surrogate_problem.py:136-137
    F1, dF1 = sp.evaluate(X[0], return_values_of=['F', 'dF'])  # 1-D input is legal here (:87-88, 154-157)
    assert F1.shape == (2,) and dF1.shape == (2, 3)
    with pytest.raises(Exception, match='Input dimension'):
        sp.evaluate(np.zeros((2, 4)))
    acq.fitted = False
    with pytest.raises(AssertionError):
        sp.evaluate(X)


def test_constraints():
    sp = SurrogateProblem(Constrained(3, 2, n_constr=2), StubAcq(2, 3))
    X = np.array([[0.9, 0.0, 0.0], [0.1, 0.0, 0.0]])
    F, CV = sp.evaluate(X)  # "auto" returns F and CV with constraints (:95-98)
    assert np.allclose(CV, [[0.4], [0.0]])
    out = sp.evaluate(X, return_values_of=['F', 'G', 'feasible'], return_as_dictionary=True)
    assert out['G'].shape == (2, 2) and list(out['feasible'][:, 0]) == [False, True]
    assert sp.evaluate_constraint(X).shape == (2, 2) and sp.evaluate_constraint(X[0]).shape == (2,)
    assert SurrogateProblem(SimpleProblem(3, 2), StubAcq(2, 3)).evaluate_constraint(X) is None


@pytest.mark.gpu
def test_real_acquisition_behind_the_adapter():
    from autooed_b200 import ExpectedImprovement, GaussianProcess
    rng = np.random.default_rng(0)
    X, Y = rng.random((30, 4)), rng.random((30, 2))
    prob = SimpleProblem(4, 2, xl=-np.ones(4), xu=2 * np.ones(4))
    gp = GaussianProcess(prob, nu=3)
    gp.fit(X, Y)
    acq = ExpectedImprovement(gp)
    acq.fit(X, Y)
    sp = SurrogateProblem(prob, acq)
    Xc = rng.random((9, 4))
    F, dF, hF = sp.evaluate(Xc, return_values_of=['F', 'dF', 'hF'])
    aF, adF, ahF = acq.evaluate(Xc, dtype='continuous', gradient=True, hessian=True)
    assert np.array_equal(F, aF) and np.array_equal(dF, adF) and np.array_equal(hF, ahF)
    f1 = sp.evaluate(Xc[3])
    assert f1.shape == (2,) and np.allclose(f1, F[3], rtol=1e-12)


def test_random_selection_consumes_the_global_rng_like_the_reference():
    """`selection/random.py:15-17`: one `np.random.choice(n, size=batch, replace=False)`; registered as 'random' (factory.py:79)."""
    from autooed_b200.factory import get_selection
    from autooed_b200.selection import Random
    assert get_selection('random') is Random
    import types
    Xc = np.arange(40.0).reshape(20, 2)
    identity = types.SimpleNamespace(do=lambda A: np.asarray(A, float), undo=lambda A: np.asarray(A, float))
    np.random.seed(4)
    got = Random(types.SimpleNamespace(transformation=identity)).select(Xc, np.zeros((20, 2)), Xc[:3], np.zeros((3, 2)), 5)
    np.random.seed(4)
    assert np.array_equal(got, Xc[np.random.choice(20, size=5, replace=False)])
