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

`pareto_discover` returns what `_pareto_discover` puts on its result queue.  The second half of the module is the algorithm
around it: stochastic sampling, the performance buffer (`discovery_buffer.py`), the family-aware batch proposal and the
`ParetoDiscovery` solver class (factory name 'discovery').
"""
import numpy as np
from scipy.linalg import null_space
from scipy.optimize import minimize

from autooed_b200.solver.base import Solver
from autooed_b200.surrogate_model import lbfgsb_lockstep
from autooed_b200.utils.sampling import lhs

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


# =====================================================================================================================
# The algorithm around the inner loop: performance buffer, stochastic sampling, family-aware batch proposal, solver class
# (`pareto_discovery.py:344-641`, `utils.py:5-134`)
# =====================================================================================================================
def stochastic_sampling(xs, xl, xu, delta_p, constr_func=None):
    """Perturbs the population by random directions of one random scale per round (section 6.2.2; :529-560), same draws
    from the global numpy RNG as the reference."""
    num_target = xs.shape[0]
    xs_final = np.zeros((0, xs.shape[1]), xs.dtype)
    while xs_final.shape[0] < num_target:
        d = np.random.random(xs.shape)
        d /= np.expand_dims(np.linalg.norm(d, axis=1), axis=1)
        delta = np.random.random() * delta_p
        xs_perturb = np.clip(xs + 1.0 / (2 ** delta) * d, xl, xu)
        if constr_func is None:
            xs_final = xs_perturb
        else:
            flag = constr_func(xs_perturb) <= 0
            if np.any(flag):
                xs_final = np.vstack((xs_final, xs_perturb[flag]))
    return xs_final[:num_target]


def _greedy_hv_pick(curr_pfront, ref_point, pred_pfront, allowed, device=0):
    """One greedy round of `utils.py:50-62`: among the `allowed` candidates the first one with the largest positive
    HV(front + y) - HV(front); -1 when none is positive.  Runs as one round of the device HVI loop (`aoed_hvi_select`,
    which re-decides near-ties in that difference form)."""
    from autooed_b200 import _lib
    Yc = _lib.as_f64(pred_pfront)
    front = _lib.as_f64(np.asarray(curr_pfront, dtype=np.float64).reshape(-1, Yc.shape[1]))
    excluded = np.ones(len(Yc), dtype=np.uint8)
    excluded[np.asarray(allowed, dtype=np.int64)] = 0
    idx = np.full(1, -1, dtype=np.int64)
    _lib.require_device()
    _lib.check(_lib.load().aoed_hvi_select(device, Yc.shape[0], Yc.shape[1], _lib.ptr(Yc), front.shape[0], _lib.ptr(front),
                                           _lib.ptr(_lib.as_f64(ref_point)), 1, _lib.ptr(excluded), _lib.ptr(idx), None))
    return int(idx[0])


def propose_next_batch(curr_pfront, ref_point, pred_pfront, pred_pset, batch_size, labels, pick=_greedy_hv_pick):
    """Greedy hypervolume batch that visits every family once per cycle (`utils.py:5-77`)."""
    curr_pfront = np.array(curr_pfront, dtype=float).reshape(-1, pred_pfront.shape[1])
    labels = np.asarray(labels)
    free = np.ones(len(pred_pset), dtype=bool)        # not yet proposed
    unvisited = np.ones(len(pred_pset), dtype=bool)   # its family has not been served in this cycle
    next_batch_indices, family_lbls_next = [], []
    if len(pred_pset) < batch_size:  # the reference pads with index 0 (:44-47)
        next_batch_indices = [0] * (batch_size - len(pred_pset))
        batch_size = len(pred_pset)
    for _ in range(batch_size):
        if not unvisited.any():
            unvisited = free.copy()
        allowed = np.flatnonzero(unvisited)
        best = pick(curr_pfront, ref_point, pred_pfront, allowed)
        if best == -1:
            best = int(np.random.choice(allowed))
        free[best] = False
        curr_pfront = np.vstack([curr_pfront, pred_pfront[best]])
        next_batch_indices.append(best)
        family_lbls_next.append(labels[best])
        unvisited[labels == labels[best]] = False
    return pred_pset[next_batch_indices].copy(), pred_pfront[next_batch_indices].copy(), family_lbls_next


def propose_next_batch_without_label(curr_pfront, ref_point, pred_pfront, pred_pset, batch_size, pick=_greedy_hv_pick):
    """`utils.py:80-134`: the same greedy loop without families."""
    curr_pfront = np.array(curr_pfront, dtype=float).reshape(-1, pred_pfront.shape[1])
    free = np.ones(len(pred_pset), dtype=bool)
    next_batch_indices = []
    if len(pred_pset) < batch_size:
        next_batch_indices = [0] * (batch_size - len(pred_pset))
        batch_size = len(pred_pset)
    for _ in range(batch_size):
        allowed = np.flatnonzero(free)
        best = pick(curr_pfront, ref_point, pred_pfront, allowed)
        if best == -1:
            best = int(np.random.choice(allowed))
        free[best] = False
        curr_pfront = np.vstack([curr_pfront, pred_pfront[best]])
        next_batch_indices.append(best)
    return pred_pset[next_batch_indices].copy(), pred_pfront[next_batch_indices].copy()


class ParetoDiscoveryAlgorithm:
    """`ParetoDiscoveryAlgorithm` (:344-603) without the pymoo `Algorithm` scaffolding: `run(problem, X0, n_gen)` is
    `minimize(problem, algo, ('n_gen', n_gen))` with the given initial sampling — generation 1 initialises the buffer,
    every further generation is one round of algorithm 1 of the paper, `_finalize` builds the sparse approximation.
    The inner loop of all samples of a generation is ONE lock-step batch on the GPU (the reference forks n_process
    workers, each walking its samples one by one)."""

    def __init__(self, pop_size=None, n_cell=None, cell_size=10, buffer_origin=None, buffer_origin_constant=1e-2, delta_b=0.2,
                 label_cost=10, delta_p=10, delta_s=0.3, n_grid_sample=100, **kwargs):
        from autooed_b200.solver.discovery_buffer import get_buffer
        self._get_buffer = get_buffer
        self.pop_size = pop_size
        self.buffer = None
        self.buffer_args = {'cell_num': pop_size if n_cell is None else n_cell, 'cell_size': cell_size, 'origin': buffer_origin,
                            'origin_constant': buffer_origin_constant, 'delta_b': delta_b, 'label_cost': label_cost}
        self.delta_p, self.delta_s, self.n_grid_sample = delta_p, delta_s, n_grid_sample
        self.patch_id = 0
        self.constr_func = None
        self.pop_x = self.pop_f = None
        self.approx_set = self.approx_front = self.fam_lbls = None
        self.n_eval = 0

    def _evaluate(self, X):
        self.n_eval += len(X)
        return self.problem.evaluate(X, return_values_of=['F'])

    def _initialize(self, X0):
        pop_x = np.array(X0, dtype=float)
        pop_constr = self.problem.evaluate_constraint(pop_x)  # :440-453
        if pop_constr is not None:
            self.constr_func = self.problem.evaluate_constraint
            pop_x = pop_x[np.all(np.atleast_2d(pop_constr.T).T <= 0, axis=1)][:self.pop_size]
        pop_f = self._evaluate(pop_x)
        self.buffer = self._get_buffer(self.problem.n_obj, **self.buffer_args)
        patch_ids = np.full(self.pop_size, self.patch_id)  # :460 (pop_size ids whatever the size of the sampling)
        self.patch_id += 1
        self.buffer.insert(pop_x, pop_f, patch_ids)
        self.pop_x = self.buffer.sample(self.pop_size)
        self.pop_f = self._evaluate(self.pop_x)

    def _next(self):
        xs = stochastic_sampling(self.pop_x.copy(), self.problem.xl, self.problem.xu, self.delta_p, self.constr_func)
        bounds = np.stack([self.problem.xl, self.problem.xu])
        x_samples, patch_ids, sample_num, origin = pareto_discover(
            xs, self.problem.evaluate, bounds, self.constr_func, self.delta_s, self.buffer.origin, self.buffer.origin_constant,
            self.n_grid_sample)
        patch_ids = np.array(patch_ids) + self.patch_id  # global patch ids (:505-506)
        self.patch_id += sample_num
        new_origin = np.minimum(self.buffer.origin, origin)
        y_samples = self._evaluate(x_samples)
        new_origin = np.minimum(np.min(y_samples, axis=0), new_origin)
        self.buffer.move_origin(new_origin)
        self.buffer.insert(x_samples, y_samples, patch_ids)
        self.pop_x = self.buffer.sample(self.pop_size)
        self.pop_f = self._evaluate(self.pop_x)

    def _finalize(self):
        self.pop_x, self.pop_f = self.buffer.flattened()
        self.fam_lbls, self.approx_set, self.approx_front = self.buffer.sparse_approximation()

    def run(self, problem, X0, n_gen):
        self.problem = problem
        self._initialize(X0)
        for _ in range(n_gen - 1):
            self._next()
        self._finalize()
        return self

    def propose_next_batch(self, curr_pfront, ref_point, batch_size):
        """:562-593."""
        approx_x, approx_y, labels = self.approx_set, self.approx_front, self.fam_lbls
        if len(approx_x) >= batch_size:
            X_next, Y_next, fam = propose_next_batch(curr_pfront, ref_point, approx_y, approx_x, batch_size, labels)
            return X_next, Y_next, fam
        remain = batch_size - len(approx_x)
        buffer_xs, buffer_ys = self.buffer.flattened()
        prop_X, prop_Y = propose_next_batch_without_label(curr_pfront, ref_point, buffer_ys, buffer_xs, remain)
        return np.vstack([approx_x, prop_X]), np.vstack([approx_y, prop_Y]), list(labels) + [-1] * remain

    def get_sparse_front(self):
        return self.fam_lbls, self.approx_set, self.approx_front


class ParetoDiscovery(Solver):
    '''
    Solver based on ParetoDiscovery [Schulz et al. 2018].
    NOTE: only compatible with direct selection.
    '''
    def __init__(self, problem, n_gen=10, pop_size=100, n_process=None, **kwargs):
        super().__init__(problem)
        self.n_gen = n_gen
        self.algo = ParetoDiscoveryAlgorithm(pop_size=pop_size, **kwargs)  # n_process: one lock-step batch replaces the fork fan-out
        self.n_eval = 0

    def _solve(self, X, Y, batch_size):
        from autooed_b200.utils.pareto import find_pareto_front
        X = np.vstack([X, lhs(X.shape[1], batch_size)])  # :628
        algo = self.algo.run(self.problem, X, self.n_gen)
        Y_candidate = algo.pop_f
        device = getattr(getattr(self.problem.acquisition, 'surrogate_model', None), 'device', 0)
        curr_pfront = find_pareto_front(Y, device=device)
        ref_point = np.max(np.vstack([Y, Y_candidate]), axis=0)
        X_candidate, Y_candidate, _ = algo.propose_next_batch(curr_pfront, ref_point, batch_size)
        self.n_eval = algo.n_eval
        return X_candidate, Y_candidate
