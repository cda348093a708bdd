"""TEST INFRASTRUCTURE ONLY — CPU restatement of the multi-objective pieces of AutoOED's hot path:
Pareto filtering (`autooed/utils/pareto.py:27-62`), exact hypervolume (pymoo 0.4.1
`Hypervolume(ref_point).calc`, reached from `selection/hvi.py:22-33`) and greedy
hypervolume-improvement batch selection (`selection/hvi.py:16-47`).

Parity status: `find_pareto_front` / `check_pareto` follow the reference line by line (pure numpy
there too).  The hypervolume indicator lives in pymoo==0.4.1 (requirements_extra.txt:1), which is
NOT in /root/reference and not installable offline, and the reference holds no tests or golden
vectors for it: **parity unpinned** for the hypervolume value's rounding.  We restate the published
definition (Lebesgue measure of the region dominated by the non-dominated subset of F and bounded
by ref_point; exact, no normalisation — pymoo `Hypervolume._calc` + Fonseca/Paquete/Lopez-Ibanez
dimension sweep) with a slicing algorithm that is exact in real arithmetic, and anchor it in tests
on closed-form cases, inclusion-exclusion and Monte-Carlo checks.
"""
import numpy as np


def find_pareto_front(Y, return_index=False):
    """utils/pareto.py:27-46 (minimisation).  Visiting order = argsort(Y[:,0]) (pareto.py:35)."""
    Y = np.asarray(Y, dtype=np.float64)
    if len(Y) == 0:
        return np.array([])
    sorted_indices = np.argsort(Y.T[0])
    pareto_indices = []
    for idx in sorted_indices:
        if not np.logical_and((Y <= Y[idx]).all(axis=1), (Y < Y[idx]).any(axis=1)).any():  # pareto.py:39
            pareto_indices.append(idx)
    front = Y[pareto_indices].copy()
    return (front, pareto_indices) if return_index else front


def check_pareto(Y):
    """utils/pareto.py:49-62 — boolean non-dominated mask."""
    Y = np.asarray(Y, dtype=np.float64)
    mask = np.zeros(len(Y), dtype=bool)
    for idx in range(len(Y)):
        if not np.logical_and((Y <= Y[idx]).all(axis=1), (Y < Y[idx]).any(axis=1)).any():
            mask[idx] = True
    return mask


def pareto_rank(Y):
    """Non-dominated sorting rank (front number, 0 = non-dominated), pymoo NonDominatedSorting
    semantics: i dom j <=> all(F_i <= F_j) and any(F_i < F_j)  (SURVEY.md Appendix F).
    Fast non-dominated sort: the domination matrix once (blocked over rows to bound memory), then fronts are peeled by
    decrementing domination counts — O(n^2 m) in total instead of per front, which is what an honest CPU arm of the
    MOBO-iteration benchmark needs at pop 1000 + offspring."""
    Y = np.asarray(Y, dtype=np.float64)
    n = len(Y)
    dom = np.zeros((n, n), dtype=bool)  # dom[i, j]: i dominates j
    for lo in range(0, n, 512):
        blk = Y[lo:lo + 512, None, :]
        dom[lo:lo + 512] = np.logical_and((blk <= Y[None, :, :]).all(2), (blk < Y[None, :, :]).any(2))
    count = dom.sum(axis=0).astype(np.int64)  # how many points dominate j
    rank = np.full(n, -1, dtype=np.int64)
    r = 0
    while (rank < 0).any():
        front = (rank < 0) & (count == 0)
        rank[front] = r
        count = count - dom[front].sum(axis=0)
        r += 1
    return rank


def _hv_boxes(Q):
    """Volume of the union of boxes [0, q], q in Q (k x m), all coordinates >= 0."""
    k, m = Q.shape
    if k == 0:
        return 0.0
    if m == 1:
        return float(Q[:, 0].max())
    if m == 2:
        order = np.argsort(-Q[:, 0], kind="stable")
        hv, best1 = 0.0, 0.0
        q = Q[order]
        for i in range(k):
            if q[i, 1] > best1:
                hv += q[i, 0] * (q[i, 1] - best1)
                best1 = q[i, 1]
        return float(hv)
    order = np.argsort(-Q[:, m - 1], kind="stable")
    q = Q[order]
    hv = 0.0
    for i in range(k):
        nxt = q[i + 1, m - 1] if i + 1 < k else 0.0
        depth = q[i, m - 1] - nxt
        if depth > 0:
            hv += depth * _hv_boxes(q[: i + 1, : m - 1])
    return float(hv)


def hypervolume(F, ref_point):
    """pymoo 0.4.1 `Hypervolume(ref_point=r).calc(F)` for minimisation: `_calc` keeps only the
    first non-dominated front of F (`NonDominatedSorting().do(F, only_non_dominated_front=True)`, rows in their
    original order; equal rows do not dominate each other), the vendored indicator then discards points with any
    coordinate above r and measures the rest against r.  The pre-filter is what makes HV(P + y) - HV(P) EXACTLY zero
    for a candidate y that a front point weakly dominates."""
    F = np.atleast_2d(np.asarray(F, dtype=np.float64))
    ref = np.asarray(ref_point, dtype=np.float64)
    if F.size == 0:
        return 0.0
    dom = np.logical_and((F[:, None, :] <= F[None, :, :]).all(2), (F[:, None, :] < F[None, :, :]).any(2))  # [j, i]: j dom i
    F = F[~dom.any(axis=0)]
    F = F[(F <= ref).all(axis=1)]
    if len(F) == 0:
        return 0.0
    return _hv_boxes(ref - F)


def hvi_select(Y_candidate, Y, batch_size, rng=None):
    """selection/hvi.py:16-47 — returns the list of selected candidate indices.
    Random fallback (hvi.py:38-39) draws from `rng` (np.random module by default, like the reference)."""
    rng = np.random if rng is None else rng
    Yc = np.asarray(Y_candidate, dtype=np.float64)
    curr_pfront = find_pareto_front(Y)
    ref_point = np.max(np.vstack([Yc, Y]), axis=0)  # hvi.py:20
    mask = np.zeros(len(Yc), dtype=bool)
    chosen = []
    for _ in range(batch_size):
        curr_hv = hypervolume(curr_pfront, ref_point)
        max_contrib, max_idx = 0.0, -1
        for idx in np.flatnonzero(~mask):
            contrib = hypervolume(np.vstack([curr_pfront, Yc[idx]]), ref_point) - curr_hv
            if contrib > max_contrib:  # strict: lowest index wins ties (hvi.py:35-37)
                max_contrib, max_idx = contrib, idx
        if max_idx == -1:
            max_idx = rng.choice(np.flatnonzero(~mask))
        mask[max_idx] = True
        curr_pfront = np.vstack([curr_pfront, Yc[max_idx]])
        chosen.append(int(max_idx))
    return chosen


def hv_contributions(Y_candidate, front, ref_point):
    """Exclusive hypervolume of each candidate w.r.t. `front` computed directly:
    vol(box(y, ref)) - HV({max(y, p)}) — the quantity the CUDA kernel evaluates."""
    Yc = np.asarray(Y_candidate, dtype=np.float64)
    ref = np.asarray(ref_point, dtype=np.float64)
    front = np.atleast_2d(np.asarray(front, dtype=np.float64))
    out = np.zeros(len(Yc))
    for i, y in enumerate(Yc):
        if (y > ref).any():
            continue
        box = float(np.prod(ref - y))
        lim = np.maximum(front, y) if front.size else np.empty((0, len(y)))
        out[i] = box - hypervolume(lim, ref)
    return out


def hv_contributions_by_difference(Y_candidate, front, ref_point):
    """HV(front u {y}) - HV(front) per candidate — literally what hvi.py:28-34 evaluates (a dominated y gives exactly 0
    because the non-dominated pre-filter of the hypervolume drops it)."""
    Yc = np.asarray(Y_candidate, dtype=np.float64)
    front = np.asarray(front, dtype=np.float64).reshape(-1, Yc.shape[1])
    base = hypervolume(front, ref_point)
    return np.array([hypervolume(np.vstack([front, y]), ref_point) - base for y in Yc])
