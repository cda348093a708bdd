"""Multi-GPU sharding on real devices: world_size 1 through the CUDA entry points, and (when the box has >= 2 GPUs)
a 2-rank NCCL run whose gathered results must equal the single-GPU results bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _data(N=120, d=5, m=2, C=1000, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, m))) + 0.05 * rng.standard_normal((N, m))
    return X, Y, rng.random((C, d))


def _model(X, Y, device=0, fit=True, **kw):
    from autooed_b200 import GaussianProcess
    from autooed_b200.problem import SimpleProblem
    gp = GaussianProcess(SimpleProblem(X.shape[1], Y.shape[1]), nu=5, device=device, **kw)
    if fit:
        gp.fit(X, Y)
    return gp


def test_world1_paths_equal_direct_calls():
    from autooed_b200 import UpperConfidenceBound
    from autooed_b200.selection.hvi import hvi_select_indices
    from autooed_b200.sharding import evaluate_sharded, fit_sharded, hvi_select_sharded, pareto_mask_sharded
    from autooed_b200.utils.pareto import check_pareto
    X, Y, Xs = _data()
    gp = _model(X, Y)
    gp2 = fit_sharded(_model(X, Y, fit=False), X, Y)
    assert np.array_equal(np.stack([g.kernel_.theta for g in gp.gps]), np.stack([g.kernel_.theta for g in gp2.gps]))
    acq = UpperConfidenceBound(gp)
    acq.fit(X, Y)
    F, dF, _ = acq.evaluate(Xs, gradient=True)
    F2, dF2, hF2 = evaluate_sharded(lambda xs: acq.evaluate(xs, gradient=True), Xs)
    assert hF2 is None and np.array_equal(F, F2) and np.array_equal(dF, dF2)
    np.random.seed(1)
    a = hvi_select_indices(F, Y, 6)
    np.random.seed(1)
    b = hvi_select_sharded(F, Y, 6)
    assert list(a) == list(b)
    assert np.array_equal(pareto_mask_sharded(F), check_pareto(F))


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from autooed_b200 import UpperConfidenceBound
    from autooed_b200.sharding import evaluate_sharded, fit_sharded, hvi_select_sharded, pareto_mask_sharded
    X, Y, Xs = _data()
    gp = fit_sharded(_model(X, Y, device=rank, fit=False, n_restarts=3, seed=5), X, Y)
    acq = UpperConfidenceBound(gp)
    acq.fit(X, Y)
    F, dF, _ = evaluate_sharded(lambda xs: acq.evaluate(xs, gradient=True), Xs)
    np.random.seed(1)
    idx = hvi_select_sharded(F, Y, 6)
    mask = pareto_mask_sharded(F)
    np.savez(os.path.join(tmp, f"rank{rank}.npz"), theta=np.stack([g.kernel_.theta for g in gp.gps]), F=F, dF=dF,
             idx=idx, mask=mask, n_local=gp.fit_info["n_problems_local"])
    dist.destroy_process_group()


def test_two_rank_nccl_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from autooed_b200 import UpperConfidenceBound
    from autooed_b200.selection.hvi import hvi_select_indices
    from autooed_b200.utils.pareto import check_pareto
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    X, Y, Xs = _data()
    gp = _model(X, Y, n_restarts=3, seed=5)
    acq = UpperConfidenceBound(gp)
    acq.fit(X, Y)
    F, dF, _ = acq.evaluate(Xs, gradient=True)
    np.random.seed(1)
    idx = hvi_select_indices(F, Y, 6)
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert z["n_local"] == 4
        assert np.array_equal(z["theta"], np.stack([g.kernel_.theta for g in gp.gps]))
        assert np.array_equal(z["F"], F) and np.array_equal(z["dF"], dF)
        assert list(z["idx"]) == list(idx)
        assert np.array_equal(z["mask"], check_pareto(F))
