"""NSGA-II solver (the caller of the acquisition): host logic on CPU with an analytic problem, and one
full MOBO proposal step on the GPU."""
import numpy as np
import pytest

from autooed_b200.problem import SimpleProblem
from autooed_b200.solver.nsga2 import NSGA2Algorithm, crowding_distance, lhs
from oracle import moo_oracle as mo


class Zdt1:
    xl, xu = np.zeros(6), np.ones(6)

    def __init__(self):
        self.calls = []

    def evaluate(self, X, return_values_of=None):
        self.calls.append(len(X))
        assert X.min() >= 0.0 and X.max() <= 1.0
        f1 = X[:, 0]
        g = 1 + 9 / 5 * X[:, 1:].sum(1)
        return np.stack([f1, g * (1 - np.sqrt(f1 / g))], 1)


def test_lhs_is_a_latin_hypercube():
    np.random.seed(0)
    U = lhs(5, 40)
    assert U.shape == (40, 5) and U.min() >= 0 and U.max() <= 1
    for k in range(5):
        assert sorted(np.minimum(np.floor(U[:, k] * 40), 39).astype(int)) == list(range(40))


def test_crowding_distance():
    F = np.array([[0.0, 1.0], [0.25, 0.5], [0.5, 0.25], [1.0, 0.0]])
    d = crowding_distance(F)
    assert np.isinf(d[0]) and np.isinf(d[3])
    assert d[1] == pytest.approx(0.5 + 0.75) and d[2] == pytest.approx(0.75 + 0.5)


def test_nsga2_converges_and_keeps_invariants():
    np.random.seed(1)
    prob = Zdt1()
    alg = NSGA2Algorithm(pop_size=60, rank_fn=mo.pareto_rank)
    X0 = lhs(6, 25)  # smaller than pop_size: taken as is (generation 1 = its evaluation)
    X, F = alg.minimize(prob, X0, 60)
    assert prob.calls[0] == 25 and len(prob.calls) == 60  # n_gen evaluator calls
    assert all(c <= 60 for c in prob.calls[1:]) and alg.n_eval == sum(prob.calls)
    assert X.shape == (60, 6) and F.shape == (60, 2)
    d2 = ((X[:, None] - X[None]) ** 2).sum(-1) + np.eye(60)
    assert d2.min() > 1e-16  # no duplicates survive
    hv = mo.hypervolume(F, [1.1, 1.1])
    assert hv > 0.84  # true front: 1.21 - 1/3 = 0.8767
    assert mo.check_pareto(F).mean() > 0.9


class ConstrainedZdt1(Zdt1):
    """ZDT1 with the cheap constraint x0 + x1 >= 0.6 (G <= 0 is feasible), pymoo `Problem.evaluate` protocol."""
    n_constr = 1

    def evaluate(self, X, return_values_of=None):
        F = super().evaluate(X)
        CV = np.maximum(0.6 - X[:, 0] - X[:, 1], 0.0)[:, None]
        out = {'F': F, 'CV': CV}
        vals = tuple(out[k] for k in return_values_of)
        return vals[0] if len(vals) == 1 else vals


def test_nsga2_is_feasibility_first_on_constrained_problems():
    """pymoo's NSGA2 (the reference's solver, nsga2.py:17-30) handles constraints feasibility-first: the tournament prefers
    the smaller violation and the survival fills up with the least infeasible only when feasible individuals run out."""
    np.random.seed(2)
    alg = NSGA2Algorithm(pop_size=40, rank_fn=mo.pareto_rank)
    # survival: 3 feasible + 5 infeasible, room for 5 => all feasible first (by rank), then the two smallest violations
    alg.pop_size = 5
    F = np.array([[0.1, 0.9], [0.5, 0.5], [0.6, 0.6], [0.0, 0.0], [0.0, 0.1], [0.2, 0.0], [0.3, 0.3], [0.05, 0.05]])
    cv = np.array([0.0, 0.0, 0.0, 0.5, 0.1, 0.3, 0.2, 0.9])
    X = np.arange(8, dtype=float)[:, None]
    Xs, Fs, rank, crowd, cvs = alg._survive(X, F, cv)
    assert list(Xs[:, 0]) == [0.0, 1.0, 2.0, 4.0, 6.0] and list(cvs) == [0.0, 0.0, 0.0, 0.1, 0.2]
    assert list(rank) == [0, 0, 1, 2, 2]
    # tournament: an infeasible contestant never beats a feasible one, whatever its rank
    np.random.seed(3)
    picks = alg._tournament(np.array([[5.0, 5.0], [0.0, 0.0]]), np.array([0.0, 9.0]), 2000, cv=np.array([0.0, 0.4]))
    assert (picks == 1).mean() < 0.3  # index 1 only wins its duels against itself
    # feasible contestants: the dominating one wins whatever the crowding distances say
    picks = alg._tournament(np.array([[1.0, 1.0], [0.5, 0.5]]), np.array([9.0, 0.0]), 2000)
    assert (picks == 0).mean() < 0.3
    # whole loop: the final population is feasible and still spreads along the constrained front
    alg = NSGA2Algorithm(pop_size=40, rank_fn=mo.pareto_rank)
    prob = ConstrainedZdt1()
    X, F = alg.minimize(prob, lhs(6, 30), 40)
    assert X.shape == (40, 6)
    assert (X[:, 0] + X[:, 1] >= 0.6 - 1e-12).all()
    assert F[:, 0].max() - F[:, 0].min() > 0.3


@pytest.mark.gpu
def test_mobo_step_end_to_end():
    from autooed_b200 import GaussianProcess, HypervolumeImprovement, Identity, init_solver
    from autooed_b200.mobo import MOBO
    np.random.seed(0)
    prob = SimpleProblem(6, 2)
    f = Zdt1()
    X = lhs(6, 20)
    Y = f.evaluate(X)
    sur = GaussianProcess(prob, nu=1)
    solver = init_solver('nsga2', {'n_gen': 30, 'pop_size': 100}, prob)
    opt = MOBO(prob, sur, Identity(sur), solver, HypervolumeImprovement(sur))
    hv0 = mo.hypervolume(Y, [1.1, 6.0])
    for _ in range(3):
        Xn = opt.optimize(X, Y, None, 5)
        assert Xn.shape == (5, 6) and Xn.min() >= 0 and Xn.max() <= 1
        X, Y = np.vstack([X, Xn]), np.vstack([Y, f.evaluate(Xn)])
    assert mo.hypervolume(Y, [1.1, 6.0]) > hv0  # the proposals improve the evaluated front
