"""world_size-2 `gloo` tests of the multi-GPU host logic (autooed_b200/sharding.py) on CPU: shard bounds, row gather,
restart-sharded fit (arg-min over ranks), sharded greedy HVI (global arg-max with the lowest-index tie-break) and the
row-sharded Pareto mask.  The per-shard compute is the CPU oracle here (test infrastructure); on a GPU box the same
functions default to the CUDA entry points (tests/test_gpu_sharding.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init(rank, world, port):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    return dist


def _run(fn, world, *args):
    port = _free_port()
    mp.spawn(fn, args=(world, port) + args, nprocs=world, join=True)


def test_shard_bounds_cover_everything():
    from autooed_b200.sharding import shard_bounds
    for n in (0, 1, 2, 7, 8, 9, 1000003):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == (-(-n // w) if n else 0)


# ---- evaluate ---------------------------------------------------------------------------------------------------
def _problem(seed=0, N=40, d=4, m=2):
    from oracle import gp_oracle as go
    rng = np.random.default_rng(seed)
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, m))) + 0.05 * rng.standard_normal((N, m))
    thetas = np.stack([np.concatenate([[0.1], rng.uniform(-0.5, 0.5, d), [-4.0]]) for _ in range(m)])
    return go.GPOracle(d, m, nu=5).fit(X, Y, thetas=thetas), X, Y


def _w_evaluate(rank, world, port, n_rows):
    dist = _init(rank, world, port)
    from autooed_b200.sharding import evaluate_sharded, shard_bounds
    orc, X, Y = _problem()
    Xs = np.random.default_rng(5).random((n_rows, X.shape[1]))
    calls = []

    def fn(xs):
        calls.append(len(xs))
        out = orc.evaluate(xs, std=True, gradient=True) if len(xs) else \
            {"F": np.empty((0, 2)), "S": np.empty((0, 2)), "dF": np.empty((0, 2, 4)), "dS": np.empty((0, 2, 4))}
        return {"F": out["F"], "S": out["S"], "dF": out["dF"], "dS": out["dS"], "hF": None}

    got = evaluate_sharded(fn, Xs)
    lo, hi = shard_bounds(n_rows, world, rank)
    assert calls == [hi - lo]
    ref = orc.evaluate(Xs, std=True, gradient=True) if n_rows else None
    assert got["hF"] is None
    for k in ("F", "S", "dF", "dS"):
        assert got[k].shape[0] == n_rows
        if n_rows:
            assert np.array_equal(got[k], ref[k]), k
    local, span = evaluate_sharded(fn, Xs, gather=False)
    assert span == (lo, hi) and local["F"].shape[0] == hi - lo
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [0, 1, 7, 64])
def test_evaluate_sharded_gloo(n_rows):
    _run(_w_evaluate, 2, n_rows)


# ---- fit -----------------------------------------------------------------------------------------------------------
def _make_cpu_gp(n_restarts):
    """GaussianProcess whose device calls are replaced by the CPU oracle: exercises the real problem list / lock-step
    optimiser / arg-min code without a GPU."""
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    from oracle import gp_oracle as go

    class CpuGP(GaussianProcess):
        def _set_train(self, Xn, Yn):
            self._Xn, self._Yn = np.asarray(Xn, float), np.asarray(Yn, float)

        def log_marginal_likelihood_batch(self, obj_idx, thetas, eval_gradient=True):
            vals = [go.lml_and_grad(th, self._Xn, self._Yn[:, o], self.nu) for o, th in zip(obj_idx, thetas)]
            return np.array([v[0] for v in vals]), np.stack([v[1] for v in vals])

        def _finalize(self, thetas):
            self.final_thetas = np.array(thetas)

    return CpuGP(SimpleProblem(3, 2), nu=3, n_restarts=n_restarts, seed=7)


def _fit_data():
    rng = np.random.default_rng(1)
    X = rng.random((25, 3))
    Y = np.sin(X @ rng.standard_normal((3, 2))) + 0.1 * rng.standard_normal((25, 2))
    return X, Y


def _w_fit(rank, world, port, n_restarts, expect):
    dist = _init(rank, world, port)
    from autooed_b200.sharding import fit_sharded
    gp = _make_cpu_gp(n_restarts)
    if rank == 1:
        gp._rng = np.random.RandomState(12345)  # a diverging local RNG must not matter: rank 0's problem list is broadcast
    X, Y = _fit_data()
    fit_sharded(gp, X, Y)
    assert gp.fitted and gp.fit_info["world"] == world
    assert gp.fit_info["n_problems_total"] == 2 * (1 + n_restarts)
    assert gp.fit_info["n_problems_local"] == len(range(rank, 2 * (1 + n_restarts), world))
    assert np.allclose(gp.final_thetas, expect, rtol=0, atol=0), (rank, gp.final_thetas, expect)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_restarts", [0, 3])
def test_fit_sharded_gloo_matches_single_process(n_restarts):
    X, Y = _fit_data()
    single = _make_cpu_gp(n_restarts)
    single.fit(X, Y)
    _run(_w_fit, 2, n_restarts, single.final_thetas)


# ---- HVI -----------------------------------------------------------------------------------------------------------
def _hvi_case(seed, C, N, m, dup=False):
    rng = np.random.default_rng(seed)
    Yc = rng.random((C, m))
    if dup:
        Yc[C // 2:] = Yc[:C - C // 2]  # exact ties across the two shards: the lower index must win
    Y = rng.random((N, m)) + 0.3
    return Yc, Y


def _w_hvi(rank, world, port, seed, C, N, m, batch, dup, expect):
    dist = _init(rank, world, port)
    from autooed_b200.sharding import hvi_select_sharded
    from oracle import moo_oracle as mo
    Yc, Y = _hvi_case(seed, C, N, m, dup)
    np.random.seed(3)
    idx = hvi_select_sharded(Yc, Y, batch, contrib_fn=mo.hv_contributions_by_difference, front_fn=mo.find_pareto_front)
    assert list(idx) == list(expect), (rank, idx, expect)
    dist.destroy_process_group()


@pytest.mark.parametrize("case", [(0, 31, 10, 2, 5, False), (1, 40, 12, 3, 6, False), (2, 20, 8, 2, 4, True),
                                  (3, 6, 30, 2, 6, False)])
def test_hvi_select_sharded_gloo_same_batch(case):
    from oracle import moo_oracle as mo
    seed, C, N, m, batch, dup = case
    Yc, Y = _hvi_case(seed, C, N, m, dup)
    np.random.seed(3)
    expect = mo.hvi_select(Yc, Y, batch)
    _run(_w_hvi, 2, seed, C, N, m, batch, dup, [int(i) for i in expect])


# ---- Pareto mask ---------------------------------------------------------------------------------------------------
def _w_pareto(rank, world, port, n, m):
    dist = _init(rank, world, port)
    from autooed_b200.sharding import pareto_mask_sharded
    from oracle import moo_oracle as mo
    Y = np.round(np.random.default_rng(n + m).random((n, m)), 1)  # coarse grid: plenty of ties and duplicates
    full = mo.check_pareto(Y)
    got = pareto_mask_sharded(Y, mask_rows_fn=lambda Yall, a, b: mo.check_pareto(Yall)[a:b])
    assert np.array_equal(got, full)
    dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(1, 2), (9, 2), (64, 3)])
def test_pareto_mask_sharded_gloo(shape):
    _run(_w_pareto, 2, *shape)
