"""Pins the ParetoDiscovery inner-loop oracle (oracle/discovery_oracle.py) to the reference's own
outputs (tests/golden/discovery_*.npz), and checks the product's host algebra (no GPU) against it on the recorded
derivatives."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import discovery_oracle as do
from oracle import gp_oracle as go

FILES = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "discovery_*.npz")))


def _load(name):
    g = dict(np.load(os.path.join(GOLDEN, name)))
    d, m = g["X"].shape[1], g["Y"].shape[1]
    orc = go.GPOracle(d, m, np.zeros(d), np.ones(d), int(g["nu"])).fit(g["X"], g["Y"], thetas=g["gp_theta"])
    return g, orc


def test_fixtures_present():
    assert len(FILES) == 3


@pytest.mark.parametrize("name", FILES)
def test_oracle_reproduces_the_reference_loop(name):
    g, orc = _load(name)
    np.random.seed(int(g["seed"]))
    out = do.pareto_discover(g["xs"], do.oracle_eval_func(orc), g["bounds"], float(g["delta_s"]), g["origin"],
                             float(g["origin_constant"]), int(g["n_grid"]))
    assert np.allclose(out["ys"], g["ys"], rtol=1e-9, atol=1e-10) and np.allclose(out["new_origin"], g["new_origin"], atol=1e-10)
    # x_opt is where an L-BFGS-B run stops (factr 1e7, pgtol 1e-5): reproducible to ~1e-6 between implementations of the model
    assert np.abs(out["x_opt"] - g["x_opt"]).max() < 5e-5
    assert np.array_equal(out["patch_ids"], g["patch_ids"])
    for i, D in enumerate(out["directions"]):
        R = g[f"directions_{i}"]
        assert D.shape == R.shape
        # a null-space basis is unique up to an orthogonal mix: compare the projectors
        assert np.abs(D @ D.T - R @ R.T).max() < 1e-3
    assert np.abs(out["x_samples"] - g["x_samples"]).max() < 5e-3


@pytest.mark.parametrize("name", FILES)
def test_product_host_algebra_on_recorded_derivatives(name):
    """`optimization_directions` / `first_order_approximation` of the product fed with the reference's own F, dF, hF at its
    own x_opt: the same random stream gives the same patches, to rounding."""
    from autooed_b200.solver import pareto_discovery as pdv
    g, _ = _load(name)
    np.random.seed(int(g["seed"]))
    m = g["Y"].shape[1]
    rows = []
    for i, x_opt in enumerate(g["x_opt"]):
        D = pdv.optimization_directions(x_opt, g["dF_opt"][i], g["hF_opt"][i], g["bounds"])
        R = g[f"directions_{i}"]
        assert D.shape == R.shape and np.abs(D @ D.T - R @ R.T).max() < 1e-6
        rows.append(pdv.first_order_approximation(x_opt, D, g["bounds"], None, int(g["n_grid"]), m))
    got = np.vstack(rows)
    assert got.shape == g["x_samples"].shape and np.abs(got - g["x_samples"]).max() < 1e-6


def test_box_constraint_helpers():
    from autooed_b200.solver import pareto_discovery as pdv
    bounds = np.array([[0.0, 0.0, -1.0], [1.0, 2.0, 1.0]])
    x = np.array([0.0, 2.0 - 1e-9, 0.3])
    act, up, lo = pdv.active_box_constraints(x, bounds)
    assert list(act) == [0, 1] and list(up) == [1] and list(lo) == [0]
    assert np.array_equal(pdv.box_constraint_jacobian(x, bounds), [[-1.0, 0, 0], [0, 1.0, 0]])
    assert pdv.box_constraint_jacobian(np.array([0.5, 1.0, 0.0]), bounds) is None
    z = pdv.reference_point(np.array([1.0, 2.0]), np.array([0.5, 1.5]), 0.3)
    f = np.array([0.5, 1.5])
    s = 2 * f / f.sum() - 1 - f / np.linalg.norm(f)
    assert np.allclose(z, np.array([1.0, 2.0]) + s / np.linalg.norm(s) * 0.3 * np.linalg.norm(f))
