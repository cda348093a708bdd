"""NSGA-II on the surrogate problem — the immediate caller of the GPU acquisition.

Reference: `autooed/mobo/solver/nsga2.py:13-30` delegates to pymoo 0.4.1 `NSGA2(pop_size)` + `minimize(..., ('n_gen', G))`.
pymoo is not part of this repository, so the algorithm is written out here with pymoo 0.4.1's defaults (SURVEY.md
Appendix F): the initial population is the given sampling as is — all of it (ranked) parents the first offspring, the cut to
pop_size happens at the first merge, as pymoo's `_initialize` runs its survival with n_survive = len(pop); binary tournament of pymoo's default type
'comp_by_dom_and_crowding' (the dominating contestant, else the larger crowding distance);
simulated binary crossover (prob 0.9, eta 15, per-variable exchange prob 0.5); polynomial mutation (eta 20, prob
1/n_var); duplicate elimination; n_offsprings = pop_size; rank-and-crowding survival; generation 1 is the evaluation of
the initial population.  Populations are NOT seed-for-seed identical with pymoo (different RNG call order); every
acquisition evaluation inside the loop is the parity-tested device call.

Every generation costs one batched `problem.evaluate` (pop_size rows) — the work the B200 path accelerates — plus
O(pop^2) host bookkeeping; non-dominated sorting runs on the GPU when `rank_fn` is the default."""
import numpy as np

from autooed_b200.solver.base import Solver
from autooed_b200.utils.sampling import lhs  # same random-number consumption as the reference's sampler


def crowding_distance(F):
    """Crowding distance of one front (larger = lonelier; boundary points get +inf)."""
    n, m = F.shape
    if n <= 2:
        return np.full(n, np.inf)
    dist = np.zeros(n)
    for k in range(m):
        order = np.argsort(F[:, k], kind='stable')
        f = F[order, k]
        span = f[-1] - f[0]
        d = np.zeros(n)
        d[0] = d[-1] = np.inf
        if span > 0:
            d[1:-1] = (f[2:] - f[:-2]) / span
        dist[order] += d
    return dist


def _default_rank(F):
    from autooed_b200.utils.pareto import pareto_rank
    return pareto_rank(F)


class NSGA2Algorithm:
    def __init__(self, pop_size=200, crossover_prob=0.9, eta_c=15.0, eta_m=20.0, rank_fn=None):
        self.pop_size = pop_size
        self.crossover_prob, self.eta_c, self.eta_m = crossover_prob, eta_c, eta_m
        self.rank_fn = rank_fn or _default_rank
        self.n_eval = 0

    # -- variation -------------------------------------------------------------------------------------------------
    def _tournament(self, F, crowd, n, cv=None):
        """Binary tournament (pymoo 0.4.1 `binary_tournament`, type 'comp_by_dom_and_crowding'): if either contestant
        violates a constraint the smaller violation wins; between feasible ones the one that DOMINATES the other
        (`Dominator.get_relation` on the objective vectors), else the larger crowding distance; a coin on ties."""
        a, b = np.random.randint(len(F), size=n), np.random.randint(len(F), size=n)
        a_lt, b_lt = (F[a] < F[b]).any(axis=1), (F[b] < F[a]).any(axis=1)
        decided = a_lt != b_lt
        better = np.where(decided, a_lt, crowd[a] > crowd[b])
        tie = ~decided & (crowd[a] == crowd[b])
        if cv is not None:
            infeas = (cv[a] > 0.0) | (cv[b] > 0.0)
            better = np.where(infeas, cv[a] < cv[b], better)
            tie = np.where(infeas, cv[a] == cv[b], tie)
        pick_a = np.where(tie, np.random.rand(n) < 0.5, better)
        return np.where(pick_a, a, b)

    def _sbx(self, P1, P2, xl, xu):
        n, d = P1.shape
        C1, C2 = P1.copy(), P2.copy()
        do_pair = np.random.rand(n) < self.crossover_prob
        do_var = (np.random.rand(n, d) < 0.5) & do_pair[:, None] & (np.abs(P1 - P2) > 1e-14)
        y1, y2 = np.minimum(P1, P2), np.maximum(P1, P2)
        delta = np.where(do_var, y2 - y1, 1.0)
        u = np.random.rand(n, d)

        def child(beta_bound):
            alpha = 2.0 - np.power(beta_bound, -(self.eta_c + 1.0))
            return np.where(u <= 1.0 / alpha, np.power(u * alpha, 1.0 / (self.eta_c + 1.0)),
                            np.power(1.0 / (2.0 - u * alpha), 1.0 / (self.eta_c + 1.0)))

        betaq1 = child(1.0 + 2.0 * (y1 - xl) / delta)
        betaq2 = child(1.0 + 2.0 * (xu - y2) / delta)
        c1 = np.clip(0.5 * ((y1 + y2) - betaq1 * delta), xl, xu)
        c2 = np.clip(0.5 * ((y1 + y2) + betaq2 * delta), xl, xu)
        swap = np.random.rand(n, d) < 0.5
        C1 = np.where(do_var, np.where(swap, c2, c1), C1)
        C2 = np.where(do_var, np.where(swap, c1, c2), C2)
        return C1, C2

    def _mutate(self, X, xl, xu):
        n, d = X.shape
        do = np.random.rand(n, d) < 1.0 / d
        span = np.where(xu > xl, xu - xl, 1.0)
        d1, d2 = (X - xl) / span, (xu - X) / span
        u = np.random.rand(n, d)
        p = 1.0 / (self.eta_m + 1.0)
        lo = np.power(2.0 * u + (1.0 - 2.0 * u) * np.power(1.0 - d1, self.eta_m + 1.0), p) - 1.0
        hi = 1.0 - np.power(2.0 * (1.0 - u) + 2.0 * (u - 0.5) * np.power(1.0 - d2, self.eta_m + 1.0), p)
        dq = np.where(u <= 0.5, lo, hi)
        return np.clip(np.where(do, X + dq * span, X), xl, xu)

    @staticmethod
    def _drop_duplicates(X, others, epsilon=1e-16):
        """pymoo `DefaultDuplicateElimination`: an offspring is discarded when its Euclidean distance to a population
        member, or to an EARLIER offspring, is <= 1e-16 (i.e. the same point).  Candidates for that are found on the first
        coordinate (sorted search) so that the check stays far below the O(n^2 d) distance matrix at pop 1000."""
        X = np.ascontiguousarray(X, dtype=float)
        ref = np.vstack([others, X]) if len(others) else X
        n0 = len(ref) - len(X)
        order = np.argsort(ref[:, 0], kind='stable')
        key = ref[order, 0]
        lo = np.searchsorted(key, X[:, 0] - epsilon, side='left')
        hi = np.searchsorted(key, X[:, 0] + epsilon, side='right')
        keep = np.ones(len(X), dtype=bool)
        for i in np.flatnonzero(hi - lo > 1):  # more than the point itself shares the first coordinate
            cand = order[lo[i]:hi[i]]
            cand = cand[cand < n0 + i]  # population members and earlier offspring only
            if len(cand) and (np.sqrt(((ref[cand] - X[i]) ** 2).sum(axis=1)) <= epsilon).any():
                keep[i] = False
        return X[keep]

    # -- survival --------------------------------------------------------------------------------------------------
    def _rank_and_crowding(self, F):
        rank = np.asarray(self.rank_fn(F))
        crowd = np.zeros(len(F))
        for r in np.unique(rank):
            idx = np.flatnonzero(rank == r)
            crowd[idx] = crowding_distance(F[idx])
        return rank, crowd

    def _survive(self, X, F, cv=None, n_survive=None):
        """Rank-and-crowding survival.  With constraints (pymoo `Survival.do`, filter_infeasible=True): the feasible
        individuals compete by rank and crowding for the `pop_size` places; places left over go to the infeasible ones
        in order of increasing violation (they carry the rank after the last feasible front and zero crowding)."""
        pop_size = self.pop_size if n_survive is None else n_survive
        if cv is None or not (cv > 0.0).any():
            rank, crowd = self._rank_and_crowding(F)
            if len(X) > pop_size:
                order = np.lexsort((-crowd, rank))[:pop_size]  # fronts in order, last front by crowding descending
                X, F, rank, crowd = X[order], F[order], rank[order], crowd[order]
                cv = None if cv is None else cv[order]
            return X, F, rank, crowd, cv
        feas, infeas = np.flatnonzero(cv <= 0.0), np.flatnonzero(cv > 0.0)
        keep, rank, crowd = np.empty(0, dtype=int), np.empty(0, dtype=int), np.empty(0)
        if len(feas):
            r, c = self._rank_and_crowding(F[feas])
            order = np.lexsort((-c, r))[:pop_size]
            keep, rank, crowd = feas[order], r[order], c[order]
        room = min(pop_size, len(X)) - len(keep)
        if room > 0:
            worst = infeas[np.argsort(cv[infeas], kind='stable')[:room]]
            base = (rank.max() + 1) if len(rank) else 0
            keep = np.concatenate([keep, worst])
            rank = np.concatenate([rank, np.full(len(worst), base, dtype=int)])
            crowd = np.concatenate([crowd, np.zeros(len(worst))])
        return X[keep], F[keep], rank, crowd, cv[keep]

    def _evaluate(self, problem, X, constrained):
        if not constrained:
            return np.asarray(problem.evaluate(X, return_values_of=['F'])), None
        F, CV = problem.evaluate(X, return_values_of=['F', 'CV'])
        return np.asarray(F), np.asarray(CV, dtype=float).reshape(-1)

    # -- driver ----------------------------------------------------------------------------------------------------
    def minimize(self, problem, X0, n_gen):
        xl, xu = np.asarray(problem.xl, dtype=float), np.asarray(problem.xu, dtype=float)
        constrained = getattr(problem, 'n_constr', 0) > 0
        X = np.array(X0, dtype=float)
        F, cv = self._evaluate(problem, X, constrained)
        self.n_eval = len(X)
        # pymoo `_initialize`: survival with n_survive = len(pop) — the whole sampling parents the first offspring
        X, F, rank, crowd, cv = self._survive(X, F, cv, n_survive=len(X))
        for _ in range(n_gen - 1):
            off = np.empty((0, X.shape[1]))
            for _attempt in range(10):  # refill after duplicate elimination, as pymoo's mating loop does
                need = self.pop_size - len(off)
                if need <= 0:
                    break
                n_pairs = (need + 1) // 2
                pa = self._tournament(F, crowd, n_pairs, cv)
                pb = self._tournament(F, crowd, n_pairs, cv)
                c1, c2 = self._sbx(X[pa], X[pb], xl, xu)
                cand = self._mutate(np.vstack([c1, c2])[:need], xl, xu)
                off = np.vstack([off, self._drop_duplicates(cand, np.vstack([X, off]))])
            if len(off) == 0:
                break
            Foff, cvoff = self._evaluate(problem, off, constrained)
            self.n_eval += len(off)
            X, F, rank, crowd, cv = self._survive(np.vstack([X, off]), np.vstack([F, Foff]),
                                                  None if cv is None else np.concatenate([cv, cvoff]))
        return X, F


def device_nsga2(acquisition, X0, xl, xu, pop_size, n_gen, seed):
    """The whole generation loop on the GPU (`aoed_nsga2_run`): no host round trip per generation."""
    import ctypes

    from autooed_b200 import _lib
    sm = acquisition.surrogate_model
    lib, h = _lib.load(), sm._ensure_handle()
    X0 = _lib.as_f64(X0)
    d, m = X0.shape[1], sm.n_obj
    param = acquisition._param()
    param = None if param is None else _lib.as_f64(param)
    Xo, Fo = np.empty((pop_size, d)), np.empty((pop_size, m))
    n_out, n_eval = ctypes.c_int64(), ctypes.c_int64()
    tail = (X0.shape[0], _lib.ptr(X0), _lib.ptr(_lib.as_f64(xl)), _lib.ptr(_lib.as_f64(xu)), pop_size, n_gen, seed, _lib.ptr(Xo),
            _lib.ptr(Fo), ctypes.byref(n_out), ctypes.byref(n_eval))
    if hasattr(acquisition, 'device_sample'):  # Thompson sampling: the loop evaluates the random-Fourier-feature sample
        W, b, theta, factor, y_mean, y_scale = acquisition.device_sample()
        _lib.check(lib.aoed_nsga2_run_rff(sm.device, d, m, W.shape[1], _lib.ptr(W), _lib.ptr(b), _lib.ptr(theta), _lib.ptr(factor),
                                          _lib.ptr(y_mean), _lib.ptr(y_scale), *tail))
    else:
        _lib.check(lib.aoed_nsga2_run(h, sm.device, d, m, acquisition.kind, _lib.ptr(param), *tail))
    return Xo[:n_out.value], Fo[:n_out.value], n_eval.value


def device_survival(F, keep, device=0):
    """One rank-and-crowding survival step on the GPU (`aoed_nsga2_survival`): (rank, crowding distance, survivor indices);
    ranks / distances are only meaningful up to the front that fills `keep` (see the header)."""
    from autooed_b200 import _lib
    F = _lib.as_f64(np.atleast_2d(F))
    n = F.shape[0]
    rank, crowd, sel = np.empty(n, dtype=np.int32), np.empty(n), np.empty(keep, dtype=np.int32)
    _lib.check(_lib.load().aoed_nsga2_survival(device, n, F.shape[1], _lib.ptr(F), keep, _lib.ptr(rank), _lib.ptr(crowd), _lib.ptr(sel)))
    return rank, crowd, sel


class NSGA2(Solver):
    '''
    Solver based on NSGA-II.
    '''
    def __init__(self, problem, n_gen=200, pop_size=200, rank_fn=None, on_device=True, **kwargs):
        super().__init__(problem)
        self.n_gen = n_gen
        self.on_device = on_device
        self.algo = NSGA2Algorithm(pop_size=pop_size, rank_fn=rank_fn)
        self.n_eval = 0

    def _device_eligible(self, n_init):
        """The device loop covers the fused GP acquisitions on unconstrained problems that fit its one-CTA sort."""
        from autooed_b200 import _lib
        acq = self.problem.acquisition
        sm = getattr(acq, 'surrogate_model', None)
        return (self.on_device and getattr(self.real_problem, 'n_constr', 0) == 0
                and (getattr(acq, 'kind', _lib.ACQ_NONE) != _lib.ACQ_NONE or hasattr(acq, 'device_sample'))
                and not getattr(acq, 'exact', False)
                and hasattr(sm, '_ensure_handle') and sm.n_var <= 64 and sm.n_obj <= 8
                and max(n_init, self.algo.pop_size) + self.algo.pop_size <= 4096)

    def _solve(self, X, Y, batch_size):
        X = np.vstack([X, lhs(X.shape[1], batch_size)])  # nsga2.py:25 (the reference samples the unit cube as well)
        if self._device_eligible(len(X)):
            seed = int(np.random.randint(0, 2 ** 31 - 1))  # the run follows the global numpy seed like the host path
            Xc, Fc, self.n_eval = device_nsga2(self.problem.acquisition, X, self.problem.xl, self.problem.xu,
                                               self.algo.pop_size, self.n_gen, seed)
            return Xc, Fc
        if self.algo.rank_fn is _default_rank:  # non-dominated sorting on the GPU the surrogate lives on
            from functools import partial
            from autooed_b200.utils.pareto import pareto_rank
            sm = getattr(self.problem.acquisition, 'surrogate_model', None)
            self.algo.rank_fn = partial(pareto_rank, device=getattr(sm, 'device', 0))
        out = self.algo.minimize(self.problem, X, self.n_gen)
        self.n_eval = self.algo.n_eval
        return out
