"""The ParetoDiscovery shell around the batched inner loop against fixtures generated from the
UNMODIFIED reference (`oracle/make_golden_buffer.py` -> tests/golden/pdshell_cases.npz): performance buffer
(`buffer.py:10-183, 306-398`), stochastic sampling (`pareto_discovery.py:529-560`), family-aware batch proposal
(`utils.py:5-134`).  Buffer and sampling are host logic (CPU tests); the proposal loop runs its hypervolume rounds on the
GPU and the solver class runs BASELINE config c3 end to end."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from autooed_b200.solver import discovery_buffer as db
from autooed_b200.solver import pareto_discovery as pdis

Z = dict(np.load(os.path.join(GOLDEN, "pdshell_cases.npz")))


def _cells_equal(buf, key):
    sizes = np.array([len(c) for c in buf.cells])
    assert np.array_equal(sizes, Z[f"{key}_sizes"])
    assert np.array_equal(buf.origin, Z[f"{key}_origin"]) and buf.sample_count == int(Z[f"{key}_count"])
    if sizes.sum() == 0:
        return
    assert np.array_equal(np.vstack([c.x for c in buf.cells if len(c)]), Z[f"{key}_x"])
    assert np.array_equal(np.vstack([c.y for c in buf.cells if len(c)]), Z[f"{key}_y"])
    assert np.array_equal(np.concatenate([c.dist for c in buf.cells if len(c)]), Z[f"{key}_dist"])
    assert np.array_equal(np.concatenate([c.pid for c in buf.cells if len(c)]), Z[f"{key}_pid"])


@pytest.mark.parametrize("tag", ["b2", "b3", "b2u"])
def test_buffer_insert_move_origin_sample_equal_reference(tag):
    """Same cells, same order inside every cell (ties included), same origin moves, same samples under the same seed."""
    m, cell_num, cell_size, d = (int(v) for v in Z[f"{tag}_cfg"])
    np.random.seed(17)
    buf = db.get_buffer(m, cell_num, cell_size=None if cell_size < 0 else cell_size, origin=None, origin_constant=1e-2,
                        delta_b=0.2, label_cost=10)
    if m == 3:
        assert np.array_equal(buf.cell_vecs, Z[f"{tag}_cell_vecs"])
    for k in range(int(Z[f"{tag}_n_ins"])):
        buf.insert(Z[f"{tag}_in{k}_X"], Z[f"{tag}_in{k}_Y"], Z[f"{tag}_in{k}_pid"])
        _cells_equal(buf, f"{tag}_after{k}")
    for n in (5, 18, 60, 400):
        np.random.seed(100 + n)
        assert np.array_equal(buf.sample(n), Z[f"{tag}_sample{n}"]), n
    fx, fy = buf.flattened()
    assert np.array_equal(fx, Z[f"{tag}_flat_x"]) and np.array_equal(fy, Z[f"{tag}_flat_y"])
    # list views with the reference's attribute names
    assert [len(c) for c in buf.buffer_x] == list(Z[f"{tag}_after{int(Z[tag + '_n_ins']) - 1}_sizes"])


def test_sparse_approximation_structure_and_cycle_limit():
    """One representative per non-empty cell, patch ids remapped in order of first appearance (buffer.py:217-226); the
    representative is the closest sample of the chosen patch, or the cell's closest sample when the cell has none of that
    patch (:263-274).  `label_cost` reaches the cut as its cycle limit (the reference passes it as pygco's fourth positional
    argument, `n_iter`): the default 0 allows no move, every cell keeps label 0.  The cut itself: tests/test_graph_cut.py."""
    for label_cost in (0, 10):
        buf = db.Buffer2D(8, cell_size=3, label_cost=label_cost)
        rng = np.random.default_rng(0)
        X, Y = rng.random((40, 3)), rng.random((40, 2)) + 0.1
        pid = rng.integers(20, 25, 40)
        buf.insert(X, Y, pid)
        labels, ax, ay = buf.sparse_approximation()
        valid = [c for c in buf.cells if len(c)]
        assert len(labels) == len(valid) == len(ax) == len(ay)
        seen = np.concatenate([c.pid for c in valid])
        assert set(seen) == set(range(len(set(seen))))  # remapped to 0 .. n_label-1
        if label_cost == 0:
            assert not np.any(labels)
        for lab, x, y, c in zip(labels, ax, ay, valid):
            hit = np.flatnonzero(c.pid == lab)
            j = hit[0] if len(hit) else 0
            assert np.array_equal(x, c.x[j]) and np.array_equal(y, c.y[j])


@pytest.mark.parametrize("tag", ["ss_free", "ss_constr"])
def test_stochastic_sampling_equals_reference(tag):
    constr = None if tag == "ss_free" else (lambda X: X[:, 0] + X[:, 1] - 1.2)
    np.random.seed(23)
    out = pdis.stochastic_sampling(Z["ss_xs"].copy(), np.zeros(5), np.ones(5), 10, constr)
    assert np.array_equal(out, Z[f"{tag}_out"])


def _oracle_pick(curr_pfront, ref_point, pred_pfront, allowed):
    """The greedy round in the reference's literal form on the CPU oracle (keeps this test GPU-free)."""
    from oracle import moo_oracle as mo
    base = mo.hypervolume(curr_pfront, ref_point)
    best, best_i = 0.0, -1
    for i in allowed:
        c = mo.hypervolume(np.vstack([curr_pfront, pred_pfront[i]]), ref_point) - base
        if c > best:
            best, best_i = c, int(i)
    return best_i


@pytest.mark.parametrize("tag", ["pnb2", "pnb3", "pnb_small"])
def test_batch_proposal_loop_equals_reference_host(tag):
    """Family cycles, masks, index-0 padding and the random fallback of utils.py:5-134 (hypervolume rounds on the oracle)."""
    args = (Z[f"{tag}_front"], Z[f"{tag}_ref"], Z[f"{tag}_pf"], Z[f"{tag}_ps"], int(Z[f"{tag}_batch"]))
    np.random.seed(3)
    Xn, Yn, fam = pdis.propose_next_batch(*args, Z[f"{tag}_labels"], pick=_oracle_pick)
    assert np.array_equal(Xn, Z[f"{tag}_X"]) and np.array_equal(Yn, Z[f"{tag}_Y"]) and np.array_equal(fam, Z[f"{tag}_fam"])
    np.random.seed(3)
    Xn, Yn = pdis.propose_next_batch_without_label(*args, pick=_oracle_pick)
    assert np.array_equal(Xn, Z[f"{tag}_X_nolabel"]) and np.array_equal(Yn, Z[f"{tag}_Y_nolabel"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["pnb2", "pnb3", "pnb_small"])
def test_batch_proposal_loop_equals_reference_gpu(tag):
    """The same fixtures with every greedy round on the device HVI loop."""
    args = (Z[f"{tag}_front"], Z[f"{tag}_ref"], Z[f"{tag}_pf"], Z[f"{tag}_ps"], int(Z[f"{tag}_batch"]))
    np.random.seed(3)
    Xn, Yn, fam = pdis.propose_next_batch(*args, Z[f"{tag}_labels"])
    assert np.array_equal(Xn, Z[f"{tag}_X"]) and np.array_equal(Yn, Z[f"{tag}_Y"]) and np.array_equal(fam, Z[f"{tag}_fam"])
    np.random.seed(3)
    Xn, Yn = pdis.propose_next_batch_without_label(*args)
    assert np.array_equal(Xn, Z[f"{tag}_X_nolabel"]) and np.array_equal(Yn, Z[f"{tag}_Y_nolabel"])


@pytest.mark.gpu
def test_discovery_solver_runs_config_c3_end_to_end():
    """BASELINE config c3 in small: ZDT3, GP + identity acquisition, ParetoDiscovery solver (gradient + Hessian calls) through
    the factory name 'discovery', direct selection of its proposal."""
    import autooed_b200 as ab
    from autooed_b200.problem import SimpleProblem
    from autooed_b200.utils.sampling import lhs
    np.random.seed(0)
    d, m = 6, 2
    prob = SimpleProblem(d, m)
    X = lhs(d, 40)
    f1 = X[:, 0]
    g = 1 + 9.0 / (d - 1) * X[:, 1:].sum(1)
    Y = np.stack([f1, g * (1 - np.sqrt(f1 / g) - f1 / g * np.sin(10 * np.pi * f1))], 1)
    sur = ab.GaussianProcess(prob, nu=5)
    sur.fit(X, Y)
    acq = ab.Identity(sur)
    acq.fit(X, Y)
    solver = ab.init_solver('discovery', {'n_gen': 3, 'pop_size': 20, 'n_grid_sample': 20}, prob)
    Xc, Yc = solver.solve(X, Y, 5, acq)
    assert Xc.shape == (5, d) and Yc.shape == (5, m) and Xc.min() >= 0 and Xc.max() <= 1
    assert np.allclose(acq.evaluate(Xc)[0], Yc, rtol=1e-9, atol=1e-12)  # the proposed values are the surrogate's
    algo = solver.algo
    assert algo.buffer.sample_count > 20 and len(algo.fam_lbls) == len(algo.approx_set) == len(algo.approx_front)
    assert solver.n_eval > 0
