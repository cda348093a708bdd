"""Lock-step L-BFGS-B: many independent bound-constrained minimisations advanced together in ONE thread, every round of
objective requests served by one batched evaluation.

The reference fits each objective's GP with `scipy.optimize.minimize(method='L-BFGS-B', jac=True)` (sklearn
`_gpr.py:_constrained_optimization`, driven from `autooed/mobo/surrogate_model/gp.py:94-101`).  scipy's routine is a
reverse-communication code: `_lbfgsb.setulb` returns whenever it wants f and g at a new point.  Driving that routine
directly — with exactly the arguments `_minimize_lbfgsb` passes — gives the same iterates bit for bit, without one Python
thread per problem and a condition-variable hand-off per evaluation (which cost about as much as the device kernel for the
small fits of a MOBO iteration).

`setulb` is a private scipy interface, so `available()` first replays a small problem against `scipy.optimize.minimize`
and compares x, f and the evaluation count exactly; any exception or difference sends the caller back to the threaded
driver.  Verified on scipy 1.15 ... 1.18 (the C port of L-BFGS-B with this `setulb` signature, `VERIFIED_SCIPY`); other
versions are not refused — the self-check decides — but `fit_info['driver']` tells which driver ran.
"""
import numpy as np

_FTOL = 2.220446049250313e-09  # scipy's defaults for method='L-BFGS-B'
_GTOL = 1e-5
_MAXCOR, _MAXLS, _MAXFUN, _MAXITER = 10, 20, 15000, 15000
_TASK_FG, _TASK_NEW_X = 3, 1

_FACTR = _FTOL / np.finfo(float).eps
_checked = None
VERIFIED_SCIPY = ((1, 15), (1, 18))  # inclusive (major, minor) range the reverse-communication driver was verified on
_setulb_fn = None


def _setulb():
    global _setulb_fn
    if _setulb_fn is None:
        from scipy.optimize import _lbfgsb
        _setulb_fn = _lbfgsb.setulb
    return _setulb_fn


class _Run:
    """State of one minimisation between two `setulb` calls (the locals of scipy's `_minimize_lbfgsb`)."""

    def __init__(self, x0, bounds):
        from scipy.optimize import _lbfgsb_py
        int_dtype = np.int64 if getattr(_lbfgsb_py, 'HAS_ILP64', False) else np.int32
        lo, hi = np.asarray(bounds, dtype=np.float64).T
        x0 = np.clip(np.asarray(x0, dtype=np.float64).ravel(), lo, hi)
        n, m = x0.size, _MAXCOR
        self.n_var = n
        self.nbd = np.zeros(n, dtype=int_dtype)
        self.low, self.up = np.zeros(n), np.zeros(n)
        for i in range(n):
            has_lo, has_hi = not np.isinf(lo[i]), not np.isinf(hi[i])
            if has_lo:
                self.low[i] = lo[i]
            if has_hi:
                self.up[i] = hi[i]
            self.nbd[i] = {(False, False): 0, (True, False): 1, (True, True): 2, (False, True): 3}[has_lo, has_hi]
        self.x = np.array(x0, dtype=np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n, dtype=np.float64)
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=int_dtype)
        self.task, self.ln_task = np.zeros(2, dtype=int_dtype), np.zeros(2, dtype=int_dtype)
        self.lsave, self.isave = np.zeros(4, dtype=int_dtype), np.zeros(44, dtype=int_dtype)
        self.dsave = np.zeros(29, dtype=np.float64)
        self.nfev = self.nit = 0
        self.done = False
        self._x_eval = self._f_eval = self._g_eval = None

    def advance(self):
        """Runs the optimiser until it asks for f, g at `self.x` (returns True) or stops (returns False)."""
        setulb = _setulb()
        while True:  # (scipy re-casts g to float64 on every pass; `feed` already guarantees that type)
            setulb(_MAXCOR, self.x, self.low, self.up, self.nbd, self.f, self.g, _FACTR, _GTOL, self.wa, self.iwa,
                   self.task, self.lsave, self.isave, self.dsave, _MAXLS, self.ln_task)
            if self.task[0] == _TASK_FG:
                # scipy's ScalarFunction re-uses f and g when the routine asks again at the point it evaluated last (it can,
                # after a line-search restart) and does not count that as an evaluation: same here
                if self._x_eval is not None and np.array_equal(self.x, self._x_eval):
                    self.f, self.g = self._f_eval, self._g_eval.copy()
                    continue
                return True
            if self.task[0] == _TASK_NEW_X:
                self.nit += 1
                if self.nit >= _MAXITER:
                    self.task[0], self.task[1] = 5, 504
                elif self.nfev > _MAXFUN:
                    self.task[0], self.task[1] = 5, 502
                continue
            self.done = True
            return False

    def feed(self, f, g):
        self.f, self.g = float(f), np.array(g, dtype=np.float64)  # own contiguous float64 copy
        self._x_eval, self._f_eval, self._g_eval = self.x.copy(), self.f, self.g.copy()
        self.nfev += 1


def minimize_lockstep(evaluate_batch, x0s, bounds):
    """Minimises len(x0s) problems; `evaluate_batch(ids, X)` returns (f, g) for the problems `ids` at the rows of X.
    Returns ([(x, f, nfev)], n_batches, n_evals)."""
    runs = [_Run(x0, bounds) for x0 in x0s]
    n_batches = n_evals = 0
    while True:
        ids = [b for b, r in enumerate(runs) if not r.done and r.advance()]
        if not ids:
            break
        f, g = evaluate_batch(ids, np.stack([runs[b].x.copy() for b in ids]))
        for i, b in enumerate(ids):
            runs[b].feed(f[i], g[i])
        n_batches += 1
        n_evals += len(ids)
    return [(r.x, r.f, r.nfev) for r in runs], n_batches, n_evals


def available():
    """True when driving `setulb` directly reproduces `scipy.optimize.minimize` exactly on a probe problem."""
    global _checked
    if _checked is None:
        try:
            from scipy.optimize import minimize
            A = np.array([[3.0, 0.5, 0.1], [0.5, 2.0, 0.3], [0.1, 0.3, 1.5]])
            c = np.array([1.0, -2.0, 0.7])

            def fg(x):
                q = A @ x - c
                return 0.5 * x @ A @ x - c @ x + 0.1 * np.sum(x ** 4), q + 0.4 * x ** 3

            bounds = np.array([[-0.25, 2.0], [-3.0, 3.0], [0.1, 0.4]])
            x0s = [np.array([1.5, 2.5, 0.2]), np.array([-0.2, -1.0, 0.4])]
            got, _, _ = minimize_lockstep(lambda ids, X: tuple(map(np.array, zip(*[fg(x) for x in X]))), x0s, bounds)
            _checked = True
            for x0, (x, f, nfev) in zip(x0s, got):
                ref = minimize(fg, x0, method='L-BFGS-B', jac=True, bounds=bounds)
                _checked = _checked and np.array_equal(ref.x, x) and ref.fun == f and ref.nfev == nfev
        except Exception:  # private interface changed: the threaded driver takes over
            _checked = False
    return _checked
