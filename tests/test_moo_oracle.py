"""Pins / anchors the multi-objective CPU oracle (oracle/moo_oracle.py).

* Pareto helpers: PINNED to the reference — `tests/golden/pareto_cases.npz` was written by the unmodified
  `autooed/utils/pareto.py` (`oracle/make_golden_moo.py`).
* Exact hypervolume: pymoo 0.4.1 is not available, so the value is anchored on closed forms, inclusion-exclusion and a
  Monte-Carlo estimate (the hypervolume is a uniquely defined number, independent of the algorithm computing it).
* Greedy HVI (`selection/hvi.py:16-47`): structural properties of the selected batch.
"""
import itertools
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import moo_oracle as mo

CASES = dict(np.load(os.path.join(GOLDEN, "pareto_cases.npz")))
NAMES = sorted({k.split("__")[0] for k in CASES})


@pytest.mark.parametrize("name", [n for n in NAMES if n != "objtype"])
def test_pareto_vs_reference(name):
    Y = CASES[f"{name}__Y"]
    front, idx = mo.find_pareto_front(Y, return_index=True)
    assert np.array_equal(mo.check_pareto(Y), CASES[f"{name}__mask"])
    # argsort is not stable in the reference: compare the index SETS and the sorted fronts
    assert sorted(int(i) for i in idx) == sorted(int(i) for i in CASES[f"{name}__idx"])
    ref = CASES[f"{name}__front"]
    assert np.array_equal(front[np.lexsort(front.T[::-1])], ref[np.lexsort(ref.T[::-1])])
    assert np.array_equal(mo.pareto_rank(Y) == 0, CASES[f"{name}__mask"])


def test_hypervolume_closed_forms():
    assert mo.hypervolume(np.array([[0.25, 0.5]]), [1.0, 1.0]) == pytest.approx(0.375)
    assert mo.hypervolume(np.array([[0.0, 0.5], [0.5, 0.0]]), [1.0, 1.0]) == pytest.approx(0.75)
    assert mo.hypervolume(np.array([[0.2, 0.3, 0.4]]), [1.0, 1.0, 1.0]) == pytest.approx(0.8 * 0.7 * 0.6)
    # a dominated point and a point outside the reference box change nothing
    F = np.array([[0.0, 0.5], [0.5, 0.0], [0.6, 0.6], [2.0, -1.0]])
    assert mo.hypervolume(F, [1.0, 1.0]) == pytest.approx(0.75)
    assert mo.hypervolume(np.empty((0, 2)), [1.0, 1.0]) == 0.0
    # points ON the reference boundary contribute zero volume
    assert mo.hypervolume(np.array([[1.0, 0.0]]), [1.0, 1.0]) == 0.0


@pytest.mark.parametrize("m", [2, 3, 4])
def test_hypervolume_inclusion_exclusion(m):
    rng = np.random.default_rng(m)
    F = rng.random((6, m))
    ref = np.full(m, 1.2)
    total = 0.0
    for k in range(1, len(F) + 1):
        for sub in itertools.combinations(range(len(F)), k):
            total += (-1) ** (k + 1) * np.prod(ref - F[list(sub)].max(axis=0))
    assert mo.hypervolume(F, ref) == pytest.approx(total, rel=1e-12)


@pytest.mark.parametrize("m", [2, 3])
def test_hypervolume_monte_carlo(m):
    rng = np.random.default_rng(10 + m)
    F = rng.random((25, m))
    ref = np.ones(m)
    U = rng.random((400000, m))
    dominated = np.zeros(len(U), dtype=bool)
    for f in F:
        dominated |= (U >= f).all(axis=1)
    assert mo.hypervolume(F, ref) == pytest.approx(dominated.mean(), abs=4e-3)


@pytest.mark.parametrize("m", [2, 3])
def test_direct_contribution_equals_difference(m):
    rng = np.random.default_rng(20 + m)
    Yc, front = rng.random((40, m)), mo.find_pareto_front(rng.random((15, m)) + 0.2)
    ref = np.full(m, 1.3)
    a = mo.hv_contributions(Yc, front, ref)
    b = mo.hv_contributions_by_difference(Yc, front, ref)
    assert np.allclose(a, b, rtol=1e-10, atol=1e-14)
    dominated = np.array([((front <= y).all(1) & (front < y).any(1)).any() for y in Yc])
    assert (b[dominated] == 0.0).all()  # exactly zero, so the strict '>' of hvi.py:35 never picks a dominated candidate


def test_greedy_hvi_properties():
    rng = np.random.default_rng(5)
    Yc, Y = rng.random((30, 2)), rng.random((10, 2)) + 0.4
    np.random.seed(0)
    idx = mo.hvi_select(Yc, Y, 8)
    assert len(idx) == len(set(idx)) == 8
    # the first pick maximises the improvement over the current front
    front = mo.find_pareto_front(Y)
    ref = np.max(np.vstack([Yc, Y]), axis=0)
    c = mo.hv_contributions_by_difference(Yc, front, ref)
    assert idx[0] == int(np.argmax(c))
    # all-dominated candidates: pure random fallback, reproducible under the seed (hvi.py:38-39)
    Yd = rng.random((12, 2)) + 2.0
    np.random.seed(1)
    a = mo.hvi_select(Yd, Y, 5)
    np.random.seed(1)
    assert a == mo.hvi_select(Yd, Y, 5)
