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


@pytest.mark.parametrize("m,C,N,batch", [(2, 200, 40, 10), (3, 500, 60, 20), (2, 30, 200, 8), (4, 100, 30, 5), (7, 40, 20, 4), (8, 30, 15, 3)])
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


def _nudge(y, rng, ulps=1):
    y = y.copy()
    k = rng.integers(len(y))
    for _ in range(ulps):
        y[k] = np.nextafter(y[k], rng.choice([-np.inf, np.inf]))
    return y


@pytest.mark.parametrize("m", [2, 3])
def test_hvi_exact_ties_and_one_ulp_near_ties(m):
    """Candidates whose contributions coincide exactly (duplicates, mirror images) or differ in the last place: the
    reference decides them by rounding of HV(front + y) - HV(front) with a strict '>' in index order (hvi.py:33-37).
    The library re-decides such rounds on the host in that very form, so the batch equals the oracle's (which evaluates
    the difference form literally) in every instance."""
    from autooed_b200.selection.hvi import hvi_select_indices
    rng = np.random.default_rng(40 + m)
    for inst in range(25):
        Y = rng.random((25, m)) * 0.6 + 0.35
        base = rng.random((12, m)) * 0.7
        rows = [base]
        rows.append(np.stack([_nudge(b, rng, ulps=int(rng.integers(1, 3))) for b in base]))  # 1-2 ulp neighbours
        rows.append(base[:4].copy())                                                            # exact duplicates
        if m == 2:
            rows.append(base[:4, ::-1].copy())                                                  # mirror images
        Yc = np.vstack(rows)
        Yc = Yc[rng.permutation(len(Yc))]
        np.random.seed(inst)
        want = mo.hvi_select(Yc, Y, 6)
        np.random.seed(inst)
        got = hvi_select_indices(Yc, Y, 6)
        assert list(got) == want, (inst, list(got), want)


def test_hvi_tie_guard_changes_nothing_without_ties(monkeypatch):
    """Generic candidates: the device-resident loop alone (guard off) and the guarded call agree with the oracle."""
    from autooed_b200.selection.hvi import hvi_select_indices
    rng = np.random.default_rng(9)
    Y = rng.random((50, 3)) + 0.2
    Yc = rng.random((800, 3))
    np.random.seed(1)
    want = mo.hvi_select(Yc, Y, 8)
    np.random.seed(1)
    assert list(hvi_select_indices(Yc, Y, 8)) == want


@pytest.mark.parametrize("m,C,batch", [(2, 100000, 12), (3, 100000, 8)])
def test_hvi_device_loop_at_1e5_candidates(m, C, batch):
    """10^5 candidates: every round of the device-resident loop equals an independent evaluation of that round (the
    one-shot contribution kernel + numpy arg-max, first maximum), and the winning contributions equal the oracle's
    HV(front + y) - HV(front) for the winner and for a sample of other candidates."""
    from autooed_b200 import _lib
    from autooed_b200.utils import pareto
    rng = np.random.default_rng(m)
    Y = rng.random((200, m)) * 0.5 + 0.5
    Yc = rng.random((C, m)) ** 0.5
    front0 = pareto.find_pareto_front(Y)
    ref = np.max(np.vstack([Yc, Y]), axis=0)
    idx = np.full(batch, -1, dtype=np.int64)
    con = np.zeros(batch)
    _lib.check(_lib.load().aoed_hvi_select(0, C, m, _lib.ptr(_lib.as_f64(Yc)), len(front0), _lib.ptr(_lib.as_f64(front0)),
                                           _lib.ptr(_lib.as_f64(ref)), batch, None, _lib.ptr(idx), _lib.ptr(con)))
    assert (idx >= 0).all() and len(set(idx)) == batch
    front, free = front0.copy(), np.ones(C, dtype=bool)
    for r in range(batch):
        c = np.where(free, pareto.hv_contributions(Yc, front, ref), 0.0)
        j = int(np.argmax(c))
        assert j == idx[r] and c[j] == con[r]
        if r % 3 == 0:
            probe = np.concatenate([[j], rng.choice(C, 20, replace=False)])
            want = mo.hv_contributions_by_difference(Yc[probe], front, ref)
            assert np.allclose(c[probe] * free[probe], want * free[probe], rtol=1e-9, atol=1e-13)
        free[j] = False
        front = np.vstack([front, Yc[j]])


@pytest.mark.parametrize("n,m", [(20000, 2), (50000, 3), (30000, 4)])
def test_pareto_filter_path_equals_direct_path(n, m, monkeypatch):
    """Above 16384 rows the mask comes from the sample-skyline filter (O(n |S0| + |R|^2)); it must equal the all-pairs
    pass bit for bit, with ties and duplicates, and on a row shard."""
    from autooed_b200 import _lib
    from autooed_b200.utils import pareto
    rng = np.random.default_rng(n + m)
    Y = rng.random((n, m))
    Y[n // 2:] = np.round(Y[n // 2:], 2)
    Y[:100] = Y[100:200]
    monkeypatch.setenv("AOED_PARETO_PATH", "direct")
    direct = pareto.check_pareto(Y)
    monkeypatch.setenv("AOED_PARETO_PATH", "filter")
    filt = pareto.check_pareto(Y)
    assert np.array_equal(direct, filt) and direct.sum() > 0
    lo, hi = n // 3, n // 3 + 5000
    out = np.zeros(hi - lo, dtype=np.uint8)
    _lib.check(_lib.load().aoed_pareto_mask_rows(0, n, m, _lib.ptr(_lib.as_f64(Y)), lo, hi, _lib.ptr(out)))
    assert np.array_equal(out.astype(bool), direct[lo:hi])
    # and the small fixtures through the filter path
    Ys = np.round(rng.random((300, m)), 1)
    assert np.array_equal(pareto.check_pareto(Ys), mo.check_pareto(Ys))


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
