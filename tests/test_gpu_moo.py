"""GPU parity of Pareto filtering and hypervolume-improvement selection against the CPU oracle
(bit-exact indices: SURVEY.md §8 a16-a17)."""
import numpy as np
import pytest

from oracle import moo_oracle as mo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m", [(1, 2), (50, 2), (300, 3), (1000, 3), (257, 4), (64, 1)])
def test_pareto_mask_and_front(n, m):
    from autooed_b200.utils import pareto
    rng = np.random.default_rng(n * 10 + m)
    Y = rng.random((n, m))
    Y[n // 2:] = np.round(Y[n // 2:], 1)  # plenty of ties and duplicates
    assert np.array_equal(pareto.check_pareto(Y), mo.check_pareto(Y))
    front, idx = pareto.find_pareto_front(Y, return_index=True)
    ofront, oidx = mo.find_pareto_front(Y, return_index=True)
    assert sorted(idx) == sorted(int(i) for i in oidx)  # argsort is not stable: compare sets
    assert np.array_equal(np.sort(front, axis=0), np.sort(ofront, axis=0))
    assert np.array_equal(pareto.pareto_rank(Y), mo.pareto_rank(Y))


@pytest.mark.parametrize("path", ["single-cta", "multi"])
def test_pareto_rank_paths(path, monkeypatch):
    """Both device paths of the non-dominated sort (one-CTA shared-memory kernel, multi-launch peeling) on a population
    with many fronts, ties and duplicates."""
    from autooed_b200.utils import pareto
    if path == "multi":
        monkeypatch.setenv("AOED_RANK_PATH", "multi")
    rng = np.random.default_rng(3)
    Y = np.round(rng.random((700, 3)), 1)
    Y[:50] = Y[50:100]
    r = pareto.pareto_rank(Y)
    assert np.array_equal(r, mo.pareto_rank(Y)) and r.max() > 5


def _peel(Y):
    n = len(Y)
    dom = (Y[:, None, :] <= Y[None, :, :]).all(-1) & (Y[:, None, :] < Y[None, :, :]).any(-1)
    cnt, want, r = dom.sum(0), np.full(n, -1), 0
    while (want < 0).any():
        front = (want < 0) & (cnt == 0)
        want[front] = r
        cnt = cnt - dom[front].sum(0)
        r += 1
    return want


@pytest.mark.parametrize("n,m", [(1500, 2), (2600, 3), (4096, 2)])
def test_pareto_rank_split_counting(n, m):
    """n above the one-kernel threshold: domination counts come from the multi-CTA counting kernel."""
    from autooed_b200.utils import pareto
    rng = np.random.default_rng(n)
    Y = np.round(rng.random((n, m)), 2)
    # front peeling on the dominance matrix: the oracle's loop is too slow at this size, so it vouches for `_peel` on a slice
    assert np.array_equal(mo.pareto_rank(Y[:400]), _peel(Y[:400]))
    assert np.array_equal(pareto.pareto_rank(Y), _peel(Y))


def test_pareto_empty():
    from autooed_b200.utils import pareto
    assert len(pareto.find_pareto_front(np.empty((0, 2)))) == 0


@pytest.mark.parametrize("m", [2, 3, 4])
def test_hv_contributions_vs_oracle(m):
    from autooed_b200.utils import pareto
    rng = np.random.default_rng(m)
    Y = rng.random((60, m)) + 0.2
    Yc = rng.random((400, m)) * 1.2
    front = mo.find_pareto_front(Y)
    ref = np.max(np.vstack([Yc, Y]), axis=0)
    got = pareto.hv_contributions(Yc, front, ref)
    want = mo.hv_contributions(Yc, front, ref)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-15)
    base = mo.hypervolume(front, ref)
    diff = np.array([mo.hypervolume(np.vstack([front, y]), ref) - base for y in Yc[:50]])
    assert np.allclose(got[:50], diff, rtol=1e-9, atol=1e-13)
    # candidates weakly dominated by a front point contribute EXACTLY zero (as pymoo's non-dominated pre-filter
    # makes the reference's HV difference vanish)
    dominated = np.array([((front <= y).all(axis=1)).any() for y in Yc])
    degenerate = (Yc >= ref).any(axis=1)  # a candidate that defines the reference point spans no volume
    assert ((got == 0) == (dominated | degenerate)).all()
    assert abs(pareto.calc_hypervolume(Y, ref) - base) < 1e-12 * max(1.0, base)


@pytest.mark.parametrize("m,C,N,batch", [(2, 200, 40, 10), (3, 500, 60, 20), (2, 30, 200, 8), (4, 100, 30, 5)])
def test_hvi_select_same_batch(m, C, N, batch):
    from autooed_b200.selection.hvi import hvi_select_indices
    rng = np.random.default_rng(100 * m + C)
    Y = rng.random((N, m)) + 0.15
    Yc = rng.random((C, m))
    Yc[C // 2:] = Yc[: C - C // 2]  # duplicates: the lowest index must win ties (hvi.py:35-37)
    np.random.seed(7)
    want = mo.hvi_select(Yc, Y, batch)
    np.random.seed(7)
    got = hvi_select_indices(Yc, Y, batch)
    assert list(got) == want


def test_hvi_random_fallback_same_seed():
    """All candidates dominated by the data: every pick is the reference's np.random.choice (hvi.py:38-39)."""
    from autooed_b200.selection.hvi import hvi_select_indices
    rng = np.random.default_rng(3)
    Y = rng.random((30, 2)) * 0.1
    Yc = rng.random((20, 2)) * 0.5 + 0.5
    np.random.seed(11)
    want = mo.hvi_select(Yc, Y, 6)
    np.random.seed(11)
    got = hvi_select_indices(Yc, Y, 6)
    assert list(got) == want


def test_selection_class_pads_and_transforms():
    from autooed_b200 import GaussianProcess, HypervolumeImprovement
    from autooed_b200.problem import SimpleProblem
    rng = np.random.default_rng(5)
    gp = GaussianProcess(SimpleProblem(4, 2))
    sel = HypervolumeImprovement(gp)
    X, Y = rng.random((20, 4)), rng.random((20, 2)) + 0.3
    Xc, Yc = rng.random((3, 4)), rng.random((3, 2))
    np.random.seed(0)
    X_next = sel.select(Xc, Yc, X, Y, 5)  # fewer candidates than the batch: padded (selection/base.py:46-50)
    assert X_next.shape == (5, 4)


def _pareto_cases():
    import os
    from conftest import GOLDEN
    z = dict(np.load(os.path.join(GOLDEN, "pareto_cases.npz")))
    return z, sorted({k.split("__")[0] for k in z})


@pytest.mark.parametrize("name", _pareto_cases()[1])
def test_pareto_vs_reference_golden(name):
    """Bit-exact non-dominated mask / index set against the unmodified reference `utils/pareto.py:27-62`."""
    from autooed_b200.utils.pareto import check_pareto, find_pareto_front, pareto_rank
    z, _ = _pareto_cases()
    Y = z[f"{name}__Y"]
    kw = {"obj_type": ["min", "max", "min"]} if name == "objtype" else {}
    assert np.array_equal(check_pareto(Y, **kw), z[f"{name}__mask"])
    front, idx = find_pareto_front(Y, return_index=True, **kw)
    assert sorted(idx) == sorted(int(i) for i in z[f"{name}__idx"])
    ref = z[f"{name}__front"]
    assert np.array_equal(front[np.lexsort(front.T[::-1])], ref[np.lexsort(ref.T[::-1])])
    if not kw:
        assert np.array_equal(pareto_rank(Y) == 0, z[f"{name}__mask"])
