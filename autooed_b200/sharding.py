"""Multi-GPU sharding of the hot path — one process per GPU, `torch.distributed` for the plumbing
(SURVEY.md §8e).  Every sub-path is embarrassingly parallel; collectives only GATHER per-shard results:

  evaluate / acquisition   candidate rows split into `world` contiguous shards, model replicated; outputs
                           all-gathered only when the caller wants one array
  fit                      the (objective x restart) L-BFGS-B problems split across ranks; (theta*, -LML*) records
                           all-gathered, arg-min per objective on every rank, final factorisation recomputed locally
  HVI selection            candidates sharded, front replicated; one (contribution, index) pair per rank and greedy round
  Pareto mask              rows sharded, columns replicated

Backend `nccl` moves the gathered arrays through device tensors over NVLink; `gloo` (CPU tests) through host tensors.
The per-shard compute is injected (`fn` arguments default to the GPU entry points) so that the host logic is testable
without a device.
"""
import numpy as np
import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_bounds(n, world_size, rank):
    """Contiguous shard [lo, hi) of n rows for `rank`: ceil(n / world) rows per rank, the tail ranks may be short or empty."""
    per = -(-n // world_size) if n > 0 else 0
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def _transport_device(group=None):
    if dist.get_backend(group) == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def all_gather_rows(local, n_total, group=None):
    """All-gathers row shards produced by `shard_bounds` into the full (n_total, ...) array on every rank."""
    ws, _ = world(group)
    local = np.ascontiguousarray(local)
    if ws == 1 or n_total == 0:
        return local
    per = -(-n_total // ws)
    tail = local.shape[1:]
    dev = _transport_device(group)
    buf = torch.zeros((per,) + tail, dtype=torch.from_numpy(local[:0]).dtype)
    if local.shape[0]:
        buf[:local.shape[0]] = torch.from_numpy(local)
    buf = buf.to(dev)
    if dev.type == 'cuda':
        out = torch.empty((ws * per,) + tail, dtype=buf.dtype, device=dev)
        dist.all_gather_into_tensor(out, buf, group=group)
    else:
        parts = [torch.empty_like(buf) for _ in range(ws)]
        dist.all_gather(parts, buf, group=group)
        out = torch.cat(parts)
    return out[:n_total].cpu().numpy()


def evaluate_sharded(fn, X, group=None, gather=True):
    """Evaluates `fn(X[lo:hi])` (a tuple / dict of row-aligned arrays, None allowed) on this rank's shard.
    gather=True: every rank returns the full arrays; gather=False: the local shard results and its (lo, hi)."""
    ws, rank = world(group)
    X = np.asarray(X)
    lo, hi = shard_bounds(X.shape[0], ws, rank)
    local = fn(X[lo:hi])
    if not gather:
        return local, (lo, hi)
    g = lambda a: None if a is None else all_gather_rows(a, X.shape[0], group)
    if isinstance(local, dict):
        return {k: g(v) for k, v in local.items()}
    if isinstance(local, (tuple, list)):
        return tuple(g(v) for v in local)
    return g(local)


def evaluate_sharded_device(gp, d_X, n_rows, acq_kind, acq_param, gradient=True, group=None, out=None):
    """Device-resident sharded acquisition evaluation with an NCCL all-gather of the results.

    `d_X` is a float64 CUDA tensor (n_rows, n_var) of NORMALISED candidates replicated on every rank (the solver's
    population); each rank evaluates the contiguous shard `shard_bounds(n_rows, world, rank)` through
    `aoed_gp_eval_device` straight into its slice of the gathered output tensors, then ONE `all_gather_into_tensor`
    per output (value: n_rows x m, gradient: n_rows x m x d) runs over NVLink.  No host staging anywhere.
    Returns (A, dA, timings) with timings = {'compute_ms', 'gather_ms'} measured with CUDA events on the current stream.
    `out` = (A, dA) re-uses gathered tensors of a previous call (their row count is world * ceil(n_rows / world))."""
    ws, rank = world(group)
    m, d = gp.n_obj, gp.n_var
    per = -(-n_rows // ws)
    dev = d_X.device
    if out is None:
        A = torch.empty((ws * per, m), dtype=torch.float64, device=dev)
        dA = torch.empty((ws * per, m, d), dtype=torch.float64, device=dev) if gradient else None
    else:
        A, dA = out
    lo, hi = shard_bounds(n_rows, ws, rank)
    a_loc = A[rank * per:(rank + 1) * per]
    da_loc = dA[rank * per:(rank + 1) * per] if gradient else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    if hi > lo:
        gp.evaluate_device(d_X[lo:hi], hi - lo, std=True, gradient=gradient, acq_kind=acq_kind, acq_param=acq_param,
                           d_A=a_loc, d_dA=da_loc, stream=stream)
    e1.record()
    if ws > 1:
        dist.all_gather_into_tensor(A, a_loc, group=group)  # in place: every rank's slice lands at its offset
        if gradient:
            dist.all_gather_into_tensor(dA, da_loc, group=group)
    e2.record()
    e2.synchronize()
    return A[:n_rows], (dA[:n_rows] if gradient else None), {'compute_ms': e0.elapsed_time(e1), 'gather_ms': e1.elapsed_time(e2)}


def fit_sharded(gp, X, Y, dtype='raw', group=None):
    """`GaussianProcess.fit` with the (objective x restart) problems distributed over the ranks.  Every rank ends
    with the same fitted model (the winning thetas are gathered, each rank factorises locally)."""
    ws, rank = world(group)
    if dtype == 'raw':
        X = gp.transformation.do(X)
    if dtype in ('raw', 'continuous'):
        gp.normalization.fit(X, Y)
        X, Y = gp.normalization.do(x=X, y=Y)
    gp._set_train(X, Y)
    problems = gp._make_problems(list(range(gp.n_obj)))
    if ws > 1:  # the random restarts must be the same list everywhere: rank 0's draw wins
        box = [problems]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        problems = box[0]
    mine = list(range(rank, len(problems), ws))  # round-robin: the starts of one objective spread over the ranks
    results = gp._run_problems([problems[i] for i in mine])
    p = gp.n_var + 2
    rec = np.full((len(problems), p + 2), np.inf)
    for i, (x, fun, nfev) in zip(mine, results):
        rec[i, :p], rec[i, p], rec[i, p + 1] = x, fun, nfev
    if ws > 1:
        dev = _transport_device(group)
        t = torch.from_numpy(rec).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)  # each row is finite on exactly one rank
        rec = t.cpu().numpy()
    full = [(rec[i, :p].copy(), float(rec[i, p]), int(rec[i, p + 1])) for i in range(len(problems))]
    best = gp._pick_best(problems, full)
    gp.fit_info = dict(gp.fit_info, n_problems_total=len(problems), n_problems_local=len(mine), world=ws)
    gp._finalize(np.stack([best[o] for o in range(gp.n_obj)]))
    gp.fitted = True
    return gp


def hvi_select_sharded(Y_candidate, Y, batch_size, contrib_fn=None, front_fn=None, group=None):
    """Greedy hypervolume-improvement batch (selection/hvi.py:16-47) with the candidates sharded: per round each rank
    scores its shard, one (contribution, global index) pair per rank is gathered, the global arg-max (lowest index on
    ties, hvi.py:35-37) is appended to every replica's front.  Returns the same indices as the single-GPU path."""
    if contrib_fn is None or front_fn is None:
        from functools import partial
        from autooed_b200.utils.pareto import find_pareto_front, hv_contributions
        dev = torch.cuda.current_device() if torch.cuda.is_available() else 0  # this rank's GPU
        contrib_fn = contrib_fn or partial(hv_contributions, device=dev)
        front_fn = front_fn or partial(find_pareto_front, device=dev)
    ws, rank = world(group)
    Yc = np.ascontiguousarray(Y_candidate, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    m = Yc.shape[1]
    lo, hi = shard_bounds(len(Yc), ws, rank)
    front = np.asarray(front_fn(Y)).reshape(-1, m)
    ref = np.max(np.vstack([Yc, Y]), axis=0)  # hvi.py:20
    free = np.ones(len(Yc), dtype=bool)
    chosen = []
    for _ in range(batch_size):
        pair = np.array([0.0, -1.0])
        if hi > lo:
            c = np.where(free[lo:hi], contrib_fn(Yc[lo:hi], front, ref), 0.0)
            j = int(np.argmax(c))  # first maximum
            if c[j] > 0.0:
                pair = np.array([c[j], float(lo + j)])
        if ws > 1:
            dev = _transport_device(group)
            mine = torch.from_numpy(pair).to(dev)
            allp = [torch.empty_like(mine) for _ in range(ws)]
            dist.all_gather(allp, mine, group=group)
            pairs = torch.stack(allp).cpu().numpy()
        else:
            pairs = pair[None]
        best_v, best_i = 0.0, -1
        for v, i in pairs:  # rank order = ascending index ranges, strict '>' keeps the lowest index
            if i >= 0 and v > best_v:
                best_v, best_i = v, int(i)
        if best_i < 0:  # no positive contribution anywhere: the reference's random pick (hvi.py:38-39), drawn on rank 0 and
            pick = [int(np.random.choice(np.flatnonzero(free)))]  # broadcast — the ranks' numpy RNG states need not agree
            if ws > 1:
                dist.broadcast_object_list(pick, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            best_i = pick[0]
        free[best_i] = False
        front = np.vstack([front, Yc[best_i]])
        chosen.append(best_i)
    return np.array(chosen, dtype=np.int64)


def pareto_mask_sharded(Y, mask_rows_fn=None, group=None):
    """Non-dominated mask with the rows sharded and the columns replicated; the mask shards are all-gathered."""
    ws, rank = world(group)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    lo, hi = shard_bounds(len(Y), ws, rank)
    if mask_rows_fn is None:
        from autooed_b200 import _lib

        def mask_rows_fn(Yall, a, b):
            out = np.zeros(b - a, dtype=np.uint8)
            _lib.require_device()
            _lib.check(_lib.load().aoed_pareto_mask_rows(torch.cuda.current_device() if torch.cuda.is_available() else 0,
                                                         Yall.shape[0], Yall.shape[1], _lib.ptr(Yall), a, b, _lib.ptr(out)))
            return out
    local = np.asarray(mask_rows_fn(Y, lo, hi), dtype=np.uint8)
    return all_gather_rows(local, len(Y), group).astype(bool)
