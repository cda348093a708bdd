"""GPU parity of the batched ParetoDiscovery inner loop (`autooed_b200/solver/pareto_discovery.py`)
against the reference-generated fixtures (`tests/golden/discovery_*.npz`): lock-step local optimisation on batched
value + gradient calls, one batched Hessian call, the reference's random stream for the expansion."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
FILES = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "discovery_*.npz")))


def _setup(name):
    import autooed_b200 as ab
    from autooed_b200.problem import SimpleProblem
    from autooed_b200.surrogate_problem import SurrogateProblem
    g = dict(np.load(os.path.join(GOLDEN, name)))
    d, m = g["X"].shape[1], g["Y"].shape[1]
    prob = SimpleProblem(d, m)
    gp = ab.GaussianProcess(prob, nu=int(g["nu"]))
    gp.fit_fixed(g["X"], g["Y"], g["gp_theta"])
    acq = ab.Identity(gp)
    acq.fit(g["X"], g["Y"])
    return g, SurrogateProblem(prob, acq)


@pytest.mark.parametrize("name", FILES)
def test_batched_loop_vs_reference(name):
    from autooed_b200.solver.pareto_discovery import pareto_discover
    g, sp = _setup(name)
    calls = []

    def eval_func(x, return_values_of):
        calls.append((np.ndim(x), len(np.atleast_2d(x)), tuple(return_values_of)))
        return sp.evaluate(x, return_values_of=return_values_of)

    np.random.seed(int(g["seed"]))
    x_samples, patch_ids, n_in, new_origin = pareto_discover(g["xs"], eval_func, g["bounds"], None, float(g["delta_s"]),
                                                             g["origin"], float(g["origin_constant"]), int(g["n_grid"]))
    assert n_in == len(g["xs"]) and np.allclose(new_origin, g["new_origin"], rtol=1e-9, atol=1e-10)
    assert list(patch_ids) == list(g["patch_ids"])
    # expansions start at x_opt (an L-BFGS-B stopping point, reproducible to ~1e-5) and follow null-space directions
    assert x_samples.shape == g["x_samples"].shape and np.abs(x_samples - g["x_samples"]).max() < 5e-3
    first = np.r_[0, np.flatnonzero(np.diff(g["patch_ids"])) + 1]
    assert np.abs(x_samples[first] - g["x_opt"]).max() < 5e-5  # each patch starts with its optimised sample
    # the population moved together: every device call was a batch, none a single row
    assert all(nd == 2 for nd, _, _ in calls)
    assert calls[0] == (2, len(g["xs"]), ('F',)) and calls[-1] == (2, len(g["xs"]), ('F', 'dF', 'hF'))
    n_rounds = len(calls) - 2
    assert 0 < n_rounds < 80 and all(rv == ('F', 'dF') for _, _, rv in calls[1:-1])


def test_constrained_problem_uses_the_per_sample_path():
    from autooed_b200.solver.pareto_discovery import pareto_discover
    g, sp = _setup(FILES[0])
    np.random.seed(1)
    out = pareto_discover(g["xs"][:3], lambda x, return_values_of: sp.evaluate(x, return_values_of=return_values_of),
                          g["bounds"], lambda x: np.atleast_2d(x).sum(axis=1) - 4.0 if np.ndim(x) == 2 else np.sum(x) - 4.0,
                          float(g["delta_s"]), g["origin"], float(g["origin_constant"]), int(g["n_grid"]))
    xs, patch, n_in, _ = out
    assert n_in == 3 and len(xs) == len(patch) and (xs.sum(axis=1) <= 4.0 + 1e-6).all()
    assert xs.min() >= 0.0 and xs.max() <= 1.0
