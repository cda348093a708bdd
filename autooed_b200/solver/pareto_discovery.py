"""Batched inner loop of the ParetoDiscovery (DGEMO) solver — SURVEY.md §8 f2.

The reference (`autooed/mobo/solver/pareto_discovery/pareto_discovery.py:22-341`) handles every stochastic sample on its
own: an L-BFGS-B run towards a reference point (`_local_optimization`), then KKT dual variables and exploration
directions from the Hessians (`_get_kkt_dual_variables`, `_get_optimization_directions`), then a first-order expansion of
the local Pareto manifold (`_first_order_approximation`) — thousands of one-row surrogate calls, fanned out over forked
processes, which a CUDA context does not survive.  Here the population moves together:

  * all L-BFGS-B runs advance in lock-step (scipy's own routine, `surrogate_model/lbfgsb_lockstep.py`) and each round of
    value + gradient requests is ONE batched acquisition call on the GPU;
  * value, gradient and Hessian at all optimised samples come from ONE batched call (the Hessian kernels' only user);
  * the small per-sample algebra (SLSQP on n_obj dual variables, null space of an (n_var + n_obj + ..)-wide matrix, grid
    sampling) stays on the host, sample by sample and in the reference's order, so the global numpy random stream is
    consumed exactly as the reference consumes it.

`pareto_discover` returns what `_pareto_discover` puts on its result queue.  The rest of the algorithm (performance buffer,
its graph-cut sparse approximation through pygco, family proposals) is not part of this module.
"""
import numpy as np
from scipy.linalg import null_space
from scipy.optimize import minimize

from autooed_b200.surrogate_model import lbfgsb_lockstep

_ACTIVE_EPS = 1e-8  # distance to a bound below which the box constraint counts as active (:151)
_ZERO_EPS = 1e-8    # entries of the null-space basis below this are numerical noise (:231)


def reference_point(y, f, delta_s):
    """Target z of the local optimisation of one sample (section 6.2.3 of the paper; :37-41)."""
    f_norm = np.linalg.norm(f)
    s = 2.0 * f / np.sum(f) - 1 - f / f_norm
    s /= np.linalg.norm(s)
    return y + s * delta_s * f_norm


def _distance_and_gradient(F, dF, z):
    diff = F - z
    dist = np.linalg.norm(diff)
    return dist, (diff / dist) @ dF


def local_optimization_batched(xs, zs, eval_func, bounds):
    """argmin_x |F(x) - z_i| inside the box for every sample i, all runs in lock-step (:43-66 without constraints)."""
    box = np.asarray(bounds, dtype=float).T
    if not lbfgsb_lockstep.available():  # private scipy interface unusable: one scipy run per sample
        return np.array([minimize(lambda x, z=z: _distance_and_gradient(*eval_func(x, return_values_of=['F', 'dF']), z), x,
                                  method='L-BFGS-B', jac=True, bounds=box).x for x, z in zip(xs, zs)])

    def evaluate_batch(ids, X):
        F, dF = eval_func(X, return_values_of=['F', 'dF'])  # one device call for all live runs
        diff = F - zs[ids]
        dist = np.sqrt(np.einsum('km,km->k', diff, diff))
        return dist, np.einsum('km,kmd->kd', diff / dist[:, None], dF)

    results, _, _ = lbfgsb_lockstep.minimize_lockstep(evaluate_batch, list(xs), box)
    return np.array([r[0] for r in results])


def _local_optimization_constrained(x, z, eval_func, bounds, constr_func):
    """SLSQP variant for problems with constraint functions, one sample at a time like the reference (:62-64)."""
    def fun(x):
        return np.linalg.norm(eval_func(x, return_values_of=['F']) - z)

    def jac(x):
        return _distance_and_gradient(*eval_func(x, return_values_of=['F', 'dF']), z)[1]

    return minimize(fun, x, method='SLSQP', jac=jac, bounds=np.asarray(bounds).T,
                    constraints=({'type': 'ineq', 'fun': lambda x: -constr_func(x)},)).x


def active_box_constraints(x, bounds):
    """Indices of the active bounds: all, the upper ones, the lower ones (:140-156)."""
    upper, lower = bounds[1] - x < _ACTIVE_EPS, x - bounds[0] < _ACTIVE_EPS
    return np.where(upper | lower)[0], np.where(upper)[0], np.where(lower)[0]


def box_constraint_jacobian(x, bounds):
    """Rows +e_k / -e_k for the active upper / lower bounds, or None when no bound is active (:159-189)."""
    active, upper, _ = active_box_constraints(x, bounds)
    if len(active) == 0:
        return None
    DG = np.zeros((len(active), len(x)))
    DG[np.arange(len(active)), active] = np.where(np.isin(active, upper), 1.0, -1.0)
    return DG


def kkt_dual_variables(n_obj, DF, DG):
    """alpha >= 0 (sum 1), beta >= 0 minimising |alpha DF + beta DG|^2 by SLSQP from a random alpha (:70-137)."""
    n_act = 0 if DG is None else len(DG)
    A = DF if DG is None else np.vstack([DF, DG])

    def residual(x):  # alpha DF + beta DG, summed in the reference's order
        return x @ DF if DG is None else x[:n_obj] @ DF + x[n_obj:] @ DG

    def fun(x):
        r = residual(x)
        return 0.5 * r @ r

    def jac(x):
        return A @ residual(x)

    ones = np.concatenate([np.ones(n_obj), np.zeros(n_act)])
    const = {'type': 'eq', 'fun': lambda x: np.sum(x[:n_obj]) - 1.0, 'jac': lambda x: ones}
    alpha0 = np.random.random(n_obj)
    x0 = np.concatenate([alpha0 / np.sum(alpha0), np.zeros(n_act)])
    res = minimize(fun, x0, method='SLSQP', jac=jac, bounds=np.array([[0.0, np.inf]] * (n_obj + n_act)), constraints=const)
    return res.x[:n_obj], res.x[n_obj:]


def optimization_directions(x_opt, DF, HF, bounds):
    """Null space of the linearised KKT system, eq. (3): directions in (alpha, beta, x) along the manifold (:192-233)."""
    n_obj, n_var = DF.shape
    DG = box_constraint_jacobian(x_opt, bounds)
    n_act = 0 if DG is None else len(DG)
    alpha, beta = kkt_dual_variables(n_obj, DF, DG)
    H = HF.T @ alpha  # box constraints have no curvature
    alpha_row = np.concatenate([np.ones(n_obj), np.zeros(n_act + n_var)])
    if n_act > 0:
        slack_rows = np.column_stack([np.zeros((n_act, n_obj + n_act)), DG])
        system = np.vstack([alpha_row, slack_rows, np.column_stack([DF.T, DG.T, H])])
    else:
        system = np.vstack([alpha_row, np.column_stack([DF.T, H])])
    directions = null_space(system)
    directions[np.abs(directions) < _ZERO_EPS] = 0.0
    return directions


def first_order_approximation(x_opt, directions, bounds, constr_func, n_grid_sample, n_obj):
    """Random grid on the tangent patch spanned by n_obj - 1 directions, clipped to the feasible box (:236-296)."""
    n_var = len(x_opt)
    lower, upper = bounds[0], bounds[1]
    x_samples = np.array([x_opt])
    d_x = directions[-n_var:]
    if np.linalg.norm(d_x) < _ZERO_EPS:
        return x_samples
    n_dir = d_x.shape[1]
    if n_dir > n_obj - 1:
        pick = np.random.choice(np.arange(n_dir), n_obj - 1)
        while np.linalg.norm(d_x[:, pick]) < _ZERO_EPS:
            pick = np.random.choice(np.arange(n_dir), n_obj - 1)
        d_x = d_x[:, pick]
    elif n_dir < n_obj - 1:
        return x_samples
    d_x = d_x / np.linalg.norm(d_x) * (upper - lower)[:, None]
    for _ in range(11):  # the reference gives up after 11 rounds without enough valid samples
        if len(x_samples) >= n_grid_sample:
            break
        steps = np.sum(d_x[None, :, :] * np.random.random((n_grid_sample, 1, n_obj - 1)), axis=-1)
        cand = x_opt[None, :] + steps
        ok = (cand <= upper).all(axis=1) & (cand >= lower).all(axis=1)
        if constr_func is not None:
            ok &= constr_func(cand) <= 0
        x_samples = np.vstack([x_samples, cand[ok]])
    return x_samples[:n_grid_sample]


def pareto_discover(xs, eval_func, bounds, constr_func, delta_s, origin, origin_constant, n_grid_sample):
    """Local optimisation + first-order expansion of a batch of samples (`_pareto_discover`, :299-341).
    `eval_func` follows `SurrogateProblem.evaluate(X, return_values_of=[...])` and must accept 2-D batches.
    Returns [x_samples_all, patch_ids, sample_num, new_origin]."""
    xs = np.atleast_2d(np.asarray(xs, dtype=float))
    bounds = np.asarray(bounds, dtype=float)
    ys = eval_func(xs, return_values_of=['F'])
    new_origin = np.minimum(origin, np.min(ys, axis=0))
    if (new_origin != origin).any():
        new_origin = new_origin - origin_constant
    fs = ys - new_origin
    zs = np.array([reference_point(y, f, delta_s) for y, f in zip(ys, fs)])
    if constr_func is None:
        x_opts = local_optimization_batched(xs, zs, eval_func, bounds)
    else:
        x_opts = np.array([_local_optimization_constrained(x, z, eval_func, bounds, constr_func) for x, z in zip(xs, zs)])
    _, DF, HF = eval_func(x_opts, return_values_of=['F', 'dF', 'hF'])  # one batched second-order call
    n_obj = ys.shape[1]
    x_samples_all, patch_ids = [], []
    for i, x_opt in enumerate(x_opts):  # host algebra in the reference's order (it draws from the global numpy RNG)
        directions = optimization_directions(x_opt, DF[i], HF[i], bounds)
        x_samples = first_order_approximation(x_opt, directions, bounds, constr_func, n_grid_sample, n_obj)
        x_samples_all.append(x_samples)
        patch_ids.extend([i] * len(x_samples))
    return [np.vstack(x_samples_all), patch_ids, len(xs), new_origin]
