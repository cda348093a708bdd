"""Gaussian-process surrogate on B200 — drop-in for the reference's
`autooed/mobo/surrogate_model/gp.py` (`GaussianProcess`, :104-253).

Same constructor, same `fit(X, Y)` / `evaluate(X, std, gradient, hessian)` / `predict`, same `.gps[i]`
attribute surface (`kernel_.theta`, `kernel_.get_params()['k1__k2__length_scale']`, `L_`, `alpha_`,
`X_train_`, `fit`), but every number is produced by libaoed_b200.so on the GPU:
  fit      -> batched log-marginal-likelihood + gradient over (objective x restart) problems, driven by the
              reference's own optimiser (scipy L-BFGS-B, gp.py:94-101) running in lock-step;
  evaluate -> fused cross-covariance / triangular-product / gradient / acquisition kernels.
"""
import ctypes
import threading

import numpy as np
from scipy.optimize import minimize

from autooed_b200 import _lib
from autooed_b200.surrogate_model.base import SurrogateModel
from autooed_b200.utils.blas import single_threaded_blas


def initial_theta(n_var):
    """log[c1=1, l=1..1, c2=1e-2]  (gp.py:126-132)."""
    return np.log(np.concatenate([[1.0], np.ones(n_var), [1e-2]]))


def theta_bounds(n_var):
    """log-bounds of gp.py:126-132: c1, l in [sqrt(1e-3), sqrt(1e3)], c2 in [e^-6, e^0]."""
    lo = np.concatenate([[np.log(np.sqrt(1e-3))], np.full(n_var, np.log(np.sqrt(1e-3))), [-6.0]])
    hi = np.concatenate([[np.log(np.sqrt(1e3))], np.full(n_var, np.log(np.sqrt(1e3))), [0.0]])
    return np.stack([lo, hi], axis=1)


class _LockstepLML:
    """Fallback driver (used when scipy's private reverse-communication routine is not usable, see lbfgsb_lockstep.py):
    runs one scipy L-BFGS-B per problem in its own thread; objective calls from all live threads are
    gathered and served by ONE batched device evaluation (`aoed_gp_lml_batch`)."""

    def __init__(self, evaluate_batch, n_prob):
        self.evaluate_batch = evaluate_batch
        self.cv = threading.Condition()
        self.pending = {}
        self.results = {}
        self.live = n_prob
        self.n_batches = 0
        self.n_evals = 0
        self.error = None

    def objective(self, b, theta):
        with self.cv:
            self.pending[b] = np.array(theta, dtype=np.float64)
            self.cv.notify_all()
            while b not in self.results and self.error is None:
                self.cv.wait()
            if self.error is not None:
                raise RuntimeError("lock-step LML evaluation failed") from self.error
            return self.results.pop(b)

    def finish(self, b):
        with self.cv:
            self.live -= 1
            self.cv.notify_all()

    def serve(self):
        with self.cv:
            while self.live > 0:
                while self.live > 0 and len(self.pending) < self.live:
                    self.cv.wait()
                if self.live == 0:
                    break
                ids = sorted(self.pending)
                thetas = np.stack([self.pending[b] for b in ids])
                self.pending.clear()
                try:
                    lml, grad = self.evaluate_batch(ids, thetas)
                except BaseException as e:  # propagate to the waiting optimiser threads
                    self.error = e
                    self.cv.notify_all()
                    raise
                self.n_batches += 1
                self.n_evals += len(ids)
                for i, b in enumerate(ids):
                    self.results[b] = (-lml[i], -grad[i])  # obj_func of _gpr.py:302-311
                self.cv.notify_all()


class _KernelView:
    """The two things other AutoOED modules read from `gp.kernel_` (ts.py:31-38, penalize.py:141)."""

    def __init__(self, theta, bounds):
        self.theta = np.array(theta)
        self.bounds = np.array(bounds)

    def get_params(self, deep=True):
        return {
            'k1__k1__constant_value': float(np.exp(self.theta[0])),
            'k1__k2__length_scale': np.exp(self.theta[1:-1]),
            'k2__constant_value': float(np.exp(self.theta[-1])),
        }


class _GPView:
    """Per-objective view with the sklearn attribute names the reference exposes through `.gps[i]`."""

    def __init__(self, owner, index):
        self._owner, self._index = owner, index
        self.kernel_ = None
        self.X_train_ = None
        self.y_train_ = None
        self.log_marginal_likelihood_value_ = None

    @property
    def L_(self):
        return self._owner._state(self._index)[0]

    @property
    def alpha_(self):
        return self._owner._state(self._index)[1]

    def fit(self, X, y):
        """Re-fit this objective alone on normalised data (what ts.py:31 does)."""
        self._owner._refit_single(self._index, X, y)
        return self

    def log_marginal_likelihood(self, theta=None, eval_gradient=False):
        if theta is None:
            return self.log_marginal_likelihood_value_
        lml, grad = self._owner.log_marginal_likelihood_batch([self._index], np.asarray(theta)[None], eval_gradient)
        return (lml[0], grad[0]) if eval_gradient else lml[0]


class GaussianProcess(SurrogateModel):
    '''
    Gaussian process (B200).
    '''
    def __init__(self, problem, nu=1, n_restarts=0, device=0, seed=None, **kwargs):
        '''
        Parameters
        ----------
        problem: autooed.problem.Problem
            The optimization problem.
        nu: int
            The parameter nu controlling the type of the Matern kernel. Choices are 1, 3, 5 and -1.
        n_restarts: int
            Extra log-uniform optimiser starts per objective (sklearn `n_restarts_optimizer` semantics,
            _gpr.py:322-340; the reference itself uses 0).
        '''
        super().__init__(problem)
        assert nu in (1, 3, 5, -1), f'Undefined nu {nu} for the Matern kernel'
        self.nu = nu
        self.n_restarts = int(n_restarts)
        self.device = int(device)
        self._rng = np.random.RandomState(seed)
        self._handle = None
        self._theta = None
        self._Xn = self._Yn = None
        self._state_cache = {}
        self.fit_info = {}
        self.gps = [_GPView(self, i) for i in range(self.n_obj)]

    # ---- device handle (created lazily: the object may be built before a fork) ----------
    def _ensure_handle(self):
        if self._handle is None:
            lib = _lib.load()
            _lib.require_device()
            h = ctypes.c_void_p()
            _lib.check(lib.aoed_gp_create(ctypes.byref(h), self.device, self.n_var, self.n_obj, self.nu))
            self._handle = h
        return self._handle

    def __del__(self):
        h, self._handle = getattr(self, '_handle', None), None
        if h is not None:
            try:
                _lib.load().aoed_gp_destroy(h)
            except Exception:
                pass

    # ---- fit -----------------------------------------------------------------------------------------------
    def log_marginal_likelihood_batch(self, obj_idx, thetas, eval_gradient=True):
        """LML (and gradient) of problem b = (objective obj_idx[b], thetas[b]) — one device batch."""
        lib, h = _lib.load(), self._ensure_handle()
        thetas = _lib.as_f64(thetas)
        obj = np.ascontiguousarray(obj_idx, dtype=np.int32)
        lml = np.empty(len(obj))
        grad = np.empty_like(thetas) if eval_gradient else None
        _lib.check(lib.aoed_gp_lml_batch(h, len(obj), _lib.ptr(obj), _lib.ptr(thetas), _lib.ptr(lml), _lib.ptr(grad)))
        self._state_cache.clear()
        return lml, grad

    def _set_train(self, Xn, Yn):
        lib, h = _lib.load(), self._ensure_handle()
        self._Xn, self._Yn = _lib.as_f64(Xn), _lib.as_f64(Yn)
        assert self._Xn.ndim == 2 and self._Xn.shape[1] == self.n_var, 'X must be (n_samples, n_var)'
        assert self._Yn.shape == (self._Xn.shape[0], self.n_obj), 'Y must be (n_samples, n_obj)'
        ys = self.normalization.y_scaler
        y_mean = _lib.as_f64(getattr(ys, 'mean_', np.zeros(self.n_obj)))
        y_scale = _lib.as_f64(getattr(ys, 'scale_', np.ones(self.n_obj)))
        _lib.check(lib.aoed_gp_set_train(h, self._Xn.shape[0], _lib.ptr(self._Xn), _lib.ptr(self._Yn),
                                         _lib.ptr(y_mean), _lib.ptr(y_scale)))

    def _make_problems(self, objectives):
        """(objective, theta0) per L-BFGS-B run: theta0 of gp.py:126-132 first, then `n_restarts` log-uniform starts
        (sklearn `n_restarts_optimizer`, _gpr.py:322-340)."""
        bounds = theta_bounds(self.n_var)
        problems = []
        for o in objectives:
            problems.append((o, initial_theta(self.n_var)))
            for _ in range(self.n_restarts):
                problems.append((o, self._rng.uniform(bounds[:, 0], bounds[:, 1])))
        return problems

    def _run_problems(self, problems):
        """One L-BFGS-B run per problem exactly as gp.py:94-101 / _gpr.py:299-340, all in lock-step on the device.
        Returns [(theta*, -lml*, nfev)] in problem order."""
        if not problems:
            self.fit_info = {'n_problems': 0, 'n_batches': 0, 'n_lml_evals': 0, 'nfev': []}
            return []
        bounds = theta_bounds(self.n_var)
        obj_of = [p[0] for p in problems]

        def evaluate_batch(ids, thetas):
            return self.log_marginal_likelihood_batch([obj_of[b] for b in ids], thetas, True)

        import os
        from autooed_b200.surrogate_model import lbfgsb_lockstep
        if os.environ.get('AOED_FIT_DRIVER', 'lockstep') != 'threads' and lbfgsb_lockstep.available():
            # scipy's own routine driven by reverse communication: same iterates, no thread hand-off per evaluation

            def neg_batch(ids, thetas):
                lml, grad = evaluate_batch(ids, thetas)
                return -lml, -grad  # obj_func of _gpr.py:302-311

            with single_threaded_blas():
                results, n_batches, n_evals = lbfgsb_lockstep.minimize_lockstep(neg_batch, [p[1] for p in problems], bounds)
            self.fit_info = {'n_problems': len(problems), 'n_batches': n_batches, 'n_lml_evals': n_evals,
                             'nfev': [r[2] for r in results], 'driver': 'lockstep'}
            return results

        pool = _LockstepLML(evaluate_batch, len(problems))
        results = [None] * len(problems)

        def worker(b):
            try:
                res = minimize(lambda th: pool.objective(b, th), problems[b][1], method="L-BFGS-B", jac=True,
                               bounds=bounds)
                results[b] = (res.x, res.fun, res.nfev)
            except BaseException as e:
                results[b] = e
            finally:
                pool.finish(b)

        threads = [threading.Thread(target=worker, args=(b,), daemon=True) for b in range(len(problems))]
        with single_threaded_blas():
            for t in threads:
                t.start()
            pool.serve()
            for t in threads:
                t.join()
        for r in results:
            if isinstance(r, BaseException):
                raise r
        self.fit_info = {'n_problems': len(problems), 'n_batches': pool.n_batches, 'n_lml_evals': pool.n_evals,
                         'nfev': [r[2] for r in results], 'driver': 'threads'}
        return results

    @staticmethod
    def _pick_best(problems, results):
        """sklearn keeps the minimum of -LML over the starts of an objective (_gpr.py:339-340); first start wins ties."""
        best = {}
        for (o, _), (x, fun, _) in zip(problems, results):
            if o not in best or fun < best[o][1]:
                best[o] = (x, fun)
        return {o: v[0] for o, v in best.items()}

    def _optimize_thetas(self, objectives):
        problems = self._make_problems(objectives)
        return self._pick_best(problems, self._run_problems(problems))

    def _finalize(self, thetas):
        lib, h = _lib.load(), self._ensure_handle()
        self._theta = _lib.as_f64(thetas)
        _lib.check(lib.aoed_gp_set_theta(h, _lib.ptr(self._theta)))
        self._state_cache.clear()
        bounds = theta_bounds(self.n_var)
        for o, g in enumerate(self.gps):
            g.kernel_ = _KernelView(self._theta[o], bounds)
            g.X_train_ = self._Xn
            g.y_train_ = self._Yn[:, o]
            lml = ctypes.c_double()
            _lib.check(lib.aoed_gp_get_state(h, o, None, None, ctypes.byref(lml)))
            g.log_marginal_likelihood_value_ = lml.value

    def _fit(self, X, Y, thetas=None):
        """X, Y normalised (base.py:31-52).  `thetas` (n_obj, n_var+2) skips the optimiser (fixed hyper-parameters)."""
        self._set_train(X, Y)
        if thetas is None:
            opt = self._optimize_thetas(list(range(self.n_obj)))
            thetas = np.stack([opt[o] for o in range(self.n_obj)])
        self._finalize(thetas)

    def fit_fixed(self, X, Y, thetas, dtype='raw'):
        """fit() with given hyper-parameters (used by benchmarks and parity tests)."""
        if dtype == 'raw':
            X = self.transformation.do(X)
        if dtype in ('raw', 'continuous'):
            self.normalization.fit(X, Y)
            X, Y = self.normalization.do(x=X, y=Y)
        self._fit(X, Y, thetas=np.asarray(thetas, dtype=np.float64))
        self.fitted = True

    def _refit_single(self, index, X, y):
        if self.fitted and self._theta is not None and self._Xn is not None and np.array_equal(X, self._Xn) \
                and np.array_equal(y, self._Yn[:, index]):
            return  # same data, same deterministic optimiser from the same start: the fit it would redo is the current one
        Yn = np.array(self._Yn if self._Yn is not None and len(self._Yn) == len(y) else np.zeros((len(y), self.n_obj)))
        Yn[:, index] = y
        thetas = np.array(self._theta) if self._theta is not None else np.tile(initial_theta(self.n_var), (self.n_obj, 1))
        self._set_train(X, Yn)
        thetas[index] = self._optimize_thetas([index])[index]
        self._finalize(thetas)

    def _state(self, index):
        if index not in self._state_cache:
            lib, h = _lib.load(), self._ensure_handle()
            n = self._Xn.shape[0]
            L, alpha = np.empty((n, n)), np.empty(n)
            _lib.check(lib.aoed_gp_get_state(h, index, _lib.ptr(L), _lib.ptr(alpha), None))
            self._state_cache[index] = (L, alpha)
        return self._state_cache[index]

    # ---- evaluate --------------------------------------------------------------------------------------------
    def _run(self, X, std, gradient, hessian, acq_kind=_lib.ACQ_NONE, acq_param=None, exact=False,
             want_surrogate=True):
        lib, h = _lib.load(), self._ensure_handle()
        X = _lib.as_f64(X)
        if X.ndim != 2 or X.shape[1] != self.n_var:
            raise ValueError(f'X must be 2-D with {self.n_var} columns, got shape {X.shape}')
        C, m, d = X.shape[0], self.n_obj, self.n_var
        need_std = std or acq_kind >= _lib.ACQ_UCB
        flags = (_lib.EVAL_STD if need_std else 0) | (_lib.EVAL_GRAD if gradient else 0) | \
                (_lib.EVAL_HESS if hessian else 0) | (_lib.EVAL_EXACT if exact else 0)
        mk = _lib.empty
        out = {'F': None, 'dF': None, 'hF': None, 'S': None, 'dS': None, 'hS': None, 'A': None, 'dA': None, 'hA': None}
        if want_surrogate:
            out['F'] = mk((C, m))
            if need_std:
                out['S'] = mk((C, m))
            if gradient:
                out['dF'] = mk((C, m, d))
                if need_std:
                    out['dS'] = mk((C, m, d))
            if hessian:
                out['hF'] = mk((C, m, d, d))
                if need_std:
                    out['hS'] = mk((C, m, d, d))
        if acq_kind != _lib.ACQ_NONE:
            out['A'] = mk((C, m))
            if gradient:
                out['dA'] = mk((C, m, d))
            if hessian:
                out['hA'] = mk((C, m, d, d))
        param = None if acq_param is None else _lib.as_f64(acq_param)
        _lib.check(lib.aoed_gp_eval(h, C, _lib.ptr(X), flags, acq_kind, _lib.ptr(param),
                                    _lib.ptr(out['F']), _lib.ptr(out['S']), _lib.ptr(out['dF']), _lib.ptr(out['dS']),
                                    _lib.ptr(out['hF']), _lib.ptr(out['hS']), _lib.ptr(out['A']), _lib.ptr(out['dA']),
                                    _lib.ptr(out['hA'])))
        if not std:
            out['S'] = out['dS'] = out['hS'] = None
        return out

    def _evaluate(self, X, std, gradient, hessian, exact=False):
        out = self._run(X, std, gradient, hessian, exact=exact)
        return {k: out[k] for k in ('F', 'dF', 'hF', 'S', 'dS', 'hS')}

    def evaluate_acquisition(self, X, kind, param, gradient=False, hessian=False, exact=False, dtype='continuous'):
        """Fused surrogate + acquisition evaluation: returns (F, dF, hF) of the acquisition."""
        assert self.fitted, 'Surrogate model is not fitted yet'
        if dtype == 'raw':
            X = self.transformation.do(X)
        if dtype in ('raw', 'continuous'):
            X = self.normalization.do(x=X)
        out = self._run(X, False, gradient, hessian, acq_kind=kind, acq_param=param, exact=exact,
                        want_surrogate=False)
        return out['A'], out['dA'], out['hA']

    def evaluate_device(self, d_X, n_rows, std=True, gradient=True, acq_kind=_lib.ACQ_NONE, acq_param=None,
                        exact=False, d_F=None, d_S=None, d_dF=None, d_dS=None, d_A=None, d_dA=None, stream=None):
        """Device-resident evaluation (no PCIe traffic): every `d_*` is a device pointer (int) or an object with
        `.data_ptr()` (e.g. a float64 torch tensor on this model's GPU) holding normalised candidates / receiving
        the outputs in the layouts of `evaluate`.  Enqueued on `stream` (cudaStream_t as int; None = legacy default)."""
        assert self.fitted, 'Surrogate model is not fitted yet'
        lib, h = _lib.load(), self._ensure_handle()

        def dp(t):
            if t is None:
                return None
            return ctypes.c_void_p(int(t.data_ptr()) if hasattr(t, 'data_ptr') else int(t))

        flags = (_lib.EVAL_STD if std else 0) | (_lib.EVAL_GRAD if gradient else 0) | (_lib.EVAL_EXACT if exact else 0)
        param = None if acq_param is None else _lib.as_f64(acq_param)
        _lib.check(lib.aoed_gp_eval_device(h, int(n_rows), dp(d_X), flags, acq_kind, _lib.ptr(param), dp(d_F), dp(d_S),
                                           dp(d_dF), dp(d_dS), None, None, dp(d_A), dp(d_dA), None,
                                           ctypes.c_void_p(stream) if stream else None))

    # ---- benchmark hooks -------------------------------------------------------------------------------------
    def counters(self):
        lib, h = _lib.load(), self._ensure_handle()
        n, ms, nt, fl = ctypes.c_int64(), ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
        _lib.check(lib.aoed_gp_get_counters(h, ctypes.byref(n), ctypes.byref(ms), ctypes.byref(nt), ctypes.byref(fl)))
        return {'launches': n.value, 'trmm_ms': ms.value, 'trmm_launches': nt.value, 'trmm_flops': fl.value}

    def fit_counters(self):
        lib, h = _lib.load(), self._ensure_handle()
        a, b = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(lib.aoed_gp_get_fit_counters(h, ctypes.byref(a), ctypes.byref(b)))
        return {'small': a.value, 'blocked': b.value}

    def reset_counters(self):
        _lib.check(_lib.load().aoed_gp_reset_counters(self._ensure_handle()))

    def set_profiling(self, on):
        _lib.check(_lib.load().aoed_gp_set_profiling(self._ensure_handle(), int(bool(on))))
