"""Device-resident NSGA-II (`aoed_nsga2_run`).  Its random stream (Philox) differs from numpy's, so
parity with the host solver is on invariants and solution quality, not on identical populations; every acquisition value it
reports must be exactly what the (parity-tested) acquisition call returns for that design."""
import numpy as np
import pytest

from oracle import moo_oracle as mo

pytestmark = pytest.mark.gpu


def _zdt1(X):
    f1 = X[:, 0]
    g = 1 + 9.0 / (X.shape[1] - 1) * X[:, 1:].sum(1)
    return np.stack([f1, g * (1 - np.sqrt(f1 / g))], 1)


def _setup(acq_name="identity", d=6, n=40, seed=0):
    import autooed_b200 as ab
    from autooed_b200.problem import SimpleProblem
    from autooed_b200.utils.sampling import lhs
    np.random.seed(seed)
    prob = SimpleProblem(d, 2)
    X = lhs(d, n)
    Y = _zdt1(X)
    sur = ab.GaussianProcess(prob, nu=5)
    sur.fit(X, Y)
    acq = ab.get_acquisition(acq_name)(sur)
    acq.fit(X, Y)
    return prob, sur, acq, X, Y


@pytest.mark.parametrize("acq_name", ["identity", "ucb", "ei"])
@pytest.mark.parametrize("pop,n_gen", [(100, 40), (30, 25)])
def test_invariants_and_reported_values(acq_name, pop, n_gen):
    from autooed_b200.solver import NSGA2
    prob, sur, acq, X, Y = _setup(acq_name)
    solver = NSGA2(prob, n_gen=n_gen, pop_size=pop)
    np.random.seed(5)
    Xc, Fc = solver.solve(X, Y, 8, acq)  # 40 + 8 initial rows: more than pop 30, fewer than pop 100
    assert Xc.shape == (pop, 6) and Fc.shape == (pop, 2)
    assert Xc.min() >= 0.0 and Xc.max() <= 1.0
    F2, _, _ = acq.evaluate(Xc)
    assert np.allclose(Fc, F2, rtol=1e-11, atol=1e-12)  # the values it optimised are the acquisition's own
    grid = np.round(Xc * 1e8).astype(np.int64)
    assert len({r.tobytes() for r in grid}) >= pop - 2  # duplicates are discarded (a few can enter while n_cur < pop)
    assert 48 <= solver.n_eval <= 48 + (n_gen - 1) * pop
    np.random.seed(5)
    Xd, Fd = NSGA2(prob, n_gen=n_gen, pop_size=pop).solve(X, Y, 8, acq)
    assert np.array_equal(Xc, Xd) and np.array_equal(Fc, Fd)  # reproducible under the numpy seed


def test_quality_matches_host_solver():
    from autooed_b200.solver import NSGA2
    prob, sur, acq, X, Y = _setup("identity", n=60)
    ref = np.array([1.2, 8.0])
    hv = {}
    for name, on_device in (("device", True), ("host", False)):
        vals = []
        for seed in (1, 2, 3):
            np.random.seed(seed)
            Xc, Fc = NSGA2(prob, n_gen=60, pop_size=100, on_device=on_device).solve(X, Y, 10, acq)
            vals.append(mo.hypervolume(Fc, ref))
        hv[name] = np.mean(vals)
    assert hv["device"] > 0.97 * hv["host"], hv
    # and both improve on the initial sampling by a wide margin
    F0, _, _ = acq.evaluate(X)
    assert hv["device"] > mo.hypervolume(F0, ref)


def test_first_generation_only_and_fallbacks():
    import autooed_b200 as ab
    from autooed_b200.solver import NSGA2
    prob, sur, acq, X, Y = _setup("identity")
    np.random.seed(0)
    Xc, Fc = NSGA2(prob, n_gen=1, pop_size=100).solve(X, Y, 8, acq)  # generation 1 = the sampling itself
    assert Xc.shape == (48, 6)
    ts = ab.ThompsonSampling(sur)
    np.random.seed(0)
    ts.fit(X, Y)
    s = NSGA2(prob, n_gen=30, pop_size=40)
    np.random.seed(1)
    Xt, Ft = s.solve(X, Y, 8, ts)  # Thompson sampling: the device loop evaluates the random-Fourier-feature sample
    assert Xt.shape == (40, 6) and s._device_eligible(48)
    F2, _, _ = ts.evaluate(Xt)
    assert np.allclose(Ft, F2, rtol=1e-11, atol=1e-12)
    np.random.seed(1)
    Xh, Fh = NSGA2(prob, n_gen=30, pop_size=40, on_device=False).solve(X, Y, 8, ts)  # host loop, same acquisition
    ref = np.maximum(Ft.max(0), Fh.max(0)) + 0.1
    assert mo.hypervolume(Ft, ref) > 0.9 * mo.hypervolume(Fh, ref)

    class Constrained(type(prob)):
        def evaluate_constraint(self, x):
            return np.array([x[0] - 0.9])
    cprob = Constrained(6, 2, n_constr=1)
    sc = NSGA2(cprob, n_gen=3, pop_size=20)
    Xc2, _ = sc.solve(X, Y, 4, acq)  # constraints are evaluated by the real problem on the host: host loop
    assert Xc2.shape == (20, 6) and not sc._device_eligible(44)


@pytest.mark.parametrize("n,m,keep,decimals,shift", [(400, 2, 200, 3, 0.0), (2000, 2, 1000, 6, 1.0), (2000, 2, 900, 6, 1.0),
                                                     (2000, 3, 1000, 2, 0.0), (1300, 4, 1000, 1, 0.0), (64, 1, 10, 6, 0.0),
                                                     (4096, 2, 2048, 3, 0.0), (4096, 3, 2000, 6, 1.0), (50, 2, 50, 2, 0.0)])
def test_survival_step_matches_host(n, m, keep, decimals, shift):
    """`aoed_nsga2_survival` (the rank / crowding / selection kernels of the device loop) against the host NSGA-II's
    rank-and-crowding survival — bit-exact: same survivors in the same order, same crowding distances.  Rounded
    objectives give ties and duplicate rows; a partly converged population gives one huge first front."""
    from autooed_b200.solver.nsga2 import NSGA2Algorithm, device_survival
    rng = np.random.default_rng(n * 10 + m)
    F = np.round(rng.random((n, m)), decimals)
    if m >= 2:  # half of the rows on a common trade-off curve: mutually non-dominated
        t = rng.random(n // 2)
        F[: n // 2, 0], F[: n // 2, 1] = t, 1.0 - t
        F[: n // 2, 2:] = 0.5
        F[n // 2:] += shift  # shift = 1: nothing dominates the curve, the first front alone has n / 2 members
    F[5] = F[3]
    rank, crowd, sel = device_survival(F, keep)
    algo = NSGA2Algorithm(pop_size=keep, rank_fn=lambda G: _peel(G))
    h_rank, h_crowd = algo._rank_and_crowding(F)
    order = np.lexsort((np.arange(n), -h_crowd, h_rank))[:keep]
    last = h_rank[order].max()
    peeled = h_rank <= last
    assert np.array_equal(rank[peeled], h_rank[peeled]) and (rank[~peeled] == last + 1).all()
    assert np.array_equal(crowd[peeled], h_crowd[peeled])
    assert np.array_equal(sel, order)


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


@pytest.mark.parametrize("acq_name", ["identity", "ei", "ts"])
def test_graph_replay_equals_stream_launches(acq_name, monkeypatch):
    """The loop replays a CUDA graph of two generations; AOED_NSGA2_GRAPH=0 issues the same kernels one by one.  Same seed
    => the same final population bit for bit (odd and even numbers of generations: the tail runs outside the graph)."""
    from autooed_b200.solver import NSGA2
    prob, sur, acq, X, Y = _setup(acq_name)
    for n_gen in (24, 31):
        out = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("AOED_NSGA2_GRAPH", mode)
            np.random.seed(9)
            if acq_name == "ts":
                acq.fit(X, Y)
            np.random.seed(9)
            out[mode] = NSGA2(prob, n_gen=n_gen, pop_size=64).solve(X, Y, 8, acq)
        assert np.array_equal(out["1"][0], out["0"][0]) and np.array_equal(out["1"][1], out["0"][1])


def test_generation_graph_variants_give_identical_runs(tmp_path):
    """The replayed graph of a generation runs the duplicate scan as a branch parallel to the evaluator and launches the kernels
    that follow each other directly as programmatic dependents (csrc/nsga2.cu).  Both are scheduling only: with them switched
    off (AOED_NSGA2_FORK=0, AOED_NSGA2_PDL=0: one serial stream, ordinary launches) and with plain stream launches instead of
    the graph (AOED_NSGA2_GRAPH=0) the same seed must give the same final population bit for bit — identity acquisition at
    pop 200 (one-CTA sort) and EI at pop 1100 (many-CTA counting + one-CTA peeling).  The switches are latched per
    process: child processes."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = f"""
import sys, numpy as np
sys.path[:0] = [{here!r}, {os.path.dirname(here)!r}]
import autooed_b200 as ab
from autooed_b200.problem import SimpleProblem
from autooed_b200.solver import NSGA2
rng = np.random.default_rng(1)
d, m, N = 5, 2, 40
X = rng.random((N, d)); Y = np.stack([X[:, 0], 1 + X[:, 1:].sum(1) - np.sqrt(X[:, 0])], 1)
prob = SimpleProblem(d, m)
sur = ab.GaussianProcess(prob, nu=5); sur.fit(X, Y)
out = {{}}
for name, pop in (("identity", 200), ("ei", 1100)):
    acq = ab.get_acquisition(name)(sur); acq.fit(X, Y)
    np.random.seed(3)
    solver = NSGA2(prob, n_gen=40, pop_size=pop)
    Xc, Fc = solver.solve(X, Y, 8, acq)
    out[name + "_X"], out[name + "_F"] = Xc, Fc
np.savez(sys.argv[1], **out)
"""
    runs = {}
    for tag, env in (("default", {}), ("serial", {"AOED_NSGA2_FORK": "0", "AOED_NSGA2_PDL": "0"}), ("stream", {"AOED_NSGA2_GRAPH": "0"})):
        path = str(tmp_path / f"{tag}.npz")
        r = subprocess.run([sys.executable, "-c", code, path], capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr[-2000:]
        runs[tag] = dict(np.load(path))
    for tag in ("serial", "stream"):
        for k, v in runs["default"].items():
            assert np.array_equal(runs[tag][k], v), (tag, k)
    assert len(runs["default"]["ei_X"]) > 0
