"""`bench.py`'s reference-literal CPU baseline: the unmodified reference surrogate staged under the
git-ignored baseline/_ref/ by `__graft_entry__.build()`.  CPU only; skipped where nothing is staged."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_literal_baseline_runs_and_agrees_with_the_port():
    import bench
    if not os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "autooed", "mobo", "surrogate_model", "gp.py")):
        pytest.skip("baseline/_ref is not staged (run __graft_entry__.build() where /root/reference is mounted)")
    out = bench.reference_literal(bench.WORKLOADS["c4"], n_train=150, rows=64, loop_rows=4)
    assert "unavailable" not in out, out
    for key in ("evaluate_std_batched_per_s", "evaluate_mean_gradient_batched_per_s", "ucb_value_gradient_row_loop_per_s",
                "sklearn_lml_gradient_call_s", "fit_fixed_theta_s"):
        assert out[key] > 0
    # the timed classes are the reference's: same posterior as the port (what `cpu_baseline.value` times) on the same inputs
    from oracle import gp_oracle as go
    from oracle import ref_loader as rl
    import warnings
    ns = rl.load()
    w = bench.WORKLOADS["c4"]
    X, Y, thetas = bench.make_problem(150, w["n_var"], w["n_obj"])
    model = ns.GaussianProcess(rl.FakeProblem(w["n_var"], w["n_obj"]), nu=w["nu"])
    for g, th in zip(model.gps, thetas):
        g.kernel = g.kernel.clone_with_theta(th)
        g.optimizer = None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.fit(X, Y)
    for g in model.gps:
        g._K_inv = None
    Xs = bench.candidates(16, w["n_var"], 2)
    ref = model.evaluate(Xs, std=True)
    orc = go.GPOracle(w["n_var"], w["n_obj"], nu=w["nu"]).fit(X, Y, thetas=thetas)
    got = orc.evaluate(Xs, std=True, var_form="kinv")
    assert np.allclose(got["F"], ref["F"], rtol=1e-9, atol=1e-11)
    assert np.allclose(got["S"], ref["S"], rtol=1e-6, atol=1e-9)
