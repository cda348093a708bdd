"""TEST INFRASTRUCTURE ONLY — CPU restatement of the ParetoDiscovery inner loop
(`autooed/mobo/solver/pareto_discovery/pareto_discovery.py:22-341`), one sample at a time as the reference runs it.
Pinned to the reference by `tests/golden/discovery_*.npz` (oracle/make_golden_discovery.py).  `eval_func(x, rv)` is any
callable with the pymoo `Problem.evaluate(x, return_values_of=rv)` protocol.  Nothing in the product package may import
this module."""
import numpy as np
from scipy.linalg import null_space
from scipy.optimize import minimize

EPS = 1e-8


def local_optimization(x, y, f, eval_func, bounds, delta_s):  # :22-67, unconstrained branch
    fn = np.linalg.norm(f)
    s = 2.0 * f / np.sum(f) - 1 - f / fn
    s /= np.linalg.norm(s)
    z = y + s * delta_s * np.linalg.norm(f)
    fun = lambda x: np.linalg.norm(eval_func(x, ['F']) - z)

    def jac(x):
        fx, dfx = eval_func(x, ['F', 'dF'])
        return ((fx - z) / np.linalg.norm(fx - z)) @ dfx
    return minimize(fun, x, method='L-BFGS-B', jac=jac, bounds=np.array(bounds).T).x


def active_sets(x, bounds):  # :140-156
    up, lo = bounds[1] - x < EPS, x - bounds[0] < EPS
    return np.where(np.logical_or(up, lo))[0], np.where(up)[0]


def directions_at(x_opt, eval_func, bounds):  # :70-137 and :192-233
    F, DF, HF = eval_func(x_opt, ['F', 'dF', 'hF'])
    act, up = active_sets(x_opt, bounds)
    n_obj, n_var, n_act = len(F), len(x_opt), len(act)
    DG = np.zeros((n_act, n_var))
    for r, k in enumerate(act):
        DG[r, k] = 1 if k in up else -1
    if n_act > 0:
        obj = lambda x: x[:n_obj] @ DF + x[n_obj:] @ DG
        fun = lambda x: 0.5 * obj(x) @ obj(x)
        jac = lambda x: np.vstack([DF, DG]) @ obj(x)
        con = {'type': 'eq', 'fun': lambda x: np.sum(x[:n_obj]) - 1.0,
               'jac': lambda x: np.concatenate([np.ones(n_obj), np.zeros_like(x[n_obj:])])}
    else:
        fun = lambda x: 0.5 * (x @ DF) @ (x @ DF)
        jac = lambda x: DF @ (x @ DF)
        con = {'type': 'eq', 'fun': lambda x: np.sum(x) - 1.0, 'jac': np.ones_like}
    a0 = np.random.random(n_obj)
    x0 = np.concatenate([a0 / np.sum(a0), np.zeros(n_act)])
    xo = minimize(fun, x0, method='SLSQP', jac=jac, bounds=np.array([[0.0, np.inf]] * (n_obj + n_act)), constraints=con).x
    alpha = xo[:n_obj]
    H = HF.T @ alpha
    head = np.concatenate([np.ones(n_obj), np.zeros(n_act + n_var)])
    if n_act > 0:
        M = np.vstack([head, np.column_stack([np.zeros((n_act, n_obj + n_act)), DG]), np.column_stack([DF.T, DG.T, H])])
    else:
        M = np.vstack([head, np.column_stack([DF.T, H])])
    D = null_space(M)
    D[np.abs(D) < EPS] = 0.0
    return D


def expand(x_opt, D, bounds, n_grid):  # :236-296 without constraint function
    n_var = len(x_opt)
    n_act = len(active_sets(x_opt, bounds)[0])
    n_obj = len(D) - n_var - n_act
    out = np.array([x_opt])
    dx = D[-n_var:]
    if np.linalg.norm(dx) < EPS:
        return out
    k = dx.shape[1]
    if k > n_obj - 1:
        idx = np.random.choice(np.arange(k), n_obj - 1)
        while np.linalg.norm(dx[:, idx]) < EPS:
            idx = np.random.choice(np.arange(k), n_obj - 1)
        dx = dx[:, idx]
    elif k < n_obj - 1:
        return out
    dx = dx / np.linalg.norm(dx)
    dx = dx * np.expand_dims(bounds[1] - bounds[0], axis=1)
    loops = 0
    while len(out) < n_grid:
        step = np.sum(np.expand_dims(dx, axis=0) * np.random.random((n_grid, 1, n_obj - 1)), axis=-1)
        cand = np.expand_dims(x_opt, axis=0) + step
        ok = np.logical_and((cand <= bounds[1]).all(axis=1), (cand >= bounds[0]).all(axis=1))
        out = np.vstack([out, cand[np.where(ok)[0]]])
        loops += 1
        if loops > 10:
            break
    return out[:n_grid]


def pareto_discover(xs, eval_func, bounds, delta_s, origin, origin_constant, n_grid):  # :299-341
    ys = eval_func(xs, ['F'])
    new_origin = np.minimum(origin, np.min(ys, axis=0))
    if (new_origin != origin).any():
        new_origin = new_origin - origin_constant
    fs = ys - new_origin
    x_opts, dirs, samples, patch = [], [], [], []
    for i, (x, y, f) in enumerate(zip(xs, ys, fs)):
        x_opt = local_optimization(x, y, f, eval_func, bounds, delta_s)
        D = directions_at(x_opt, eval_func, bounds)
        sm = expand(x_opt, D, bounds, n_grid)
        x_opts.append(x_opt); dirs.append(D); samples.append(sm); patch.extend([i] * len(sm))
    return dict(ys=ys, new_origin=new_origin, x_opt=np.array(x_opts), directions=dirs, x_samples=np.vstack(samples),
                patch_ids=np.array(patch))


def oracle_eval_func(orc):
    """pymoo-protocol evaluation on top of a GPOracle with the Identity acquisition (mean, its gradient and Hessian)."""
    def eval_func(x, rv):
        X = np.atleast_2d(x)
        out = orc.evaluate(X, std=False, gradient='dF' in rv, hessian='hF' in rv)
        vals = [out[k] if np.ndim(x) == 2 else out[k][0] for k in rv]
        return vals[0] if len(vals) == 1 else tuple(vals)
    return eval_func
