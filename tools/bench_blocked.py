"""Times the blocked LML + gradient path (csrc/blocked.cu) per device batch: python tools/bench_blocked.py [sizes...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autooed_b200 import GaussianProcess  # noqa: E402
from autooed_b200.problem import SimpleProblem  # noqa: E402
from oracle import gp_oracle as go  # noqa: E402


def run(N, d, m, batches, reps):
    rng = np.random.default_rng(0)
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, m))) + 0.02 * rng.standard_normal((N, m))
    Yn = (Y - Y.mean(0)) / Y.std(0)
    gp = GaussianProcess(SimpleProblem(d, m), nu=5)
    gp._set_train(X, Yn)
    lo, hi = go.theta_bounds(d).T
    for B in batches:
        obj = np.arange(B) % m
        th = rng.uniform(lo, hi, (B, d + 2)) * 0.3
        gp.log_marginal_likelihood_batch(obj, th, True)
        t0 = time.perf_counter()
        for _ in range(reps):
            lml, g = gp.log_marginal_likelihood_batch(obj, th, True)
        dt = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            gp.log_marginal_likelihood_batch(obj, th, False)
        dtv = (time.perf_counter() - t0) / reps
        flops = B * (N ** 3 + 2.0 * N * N * (d + 2) + N * N * (3 * d + 25))
        print(json.dumps({"what": "lml+grad batch (blocked)", "N": N, "d": d, "problems": B, "ms_per_batch": 1e3 * dt,
                          "ms_value_only": 1e3 * dtv, "tflops": flops / dt * 1e-12, "finite": bool(np.isfinite(lml).all())}), flush=True)
    th = rng.uniform(lo, hi, (m, d + 2)) * 0.3
    gp._finalize(th)
    t0 = time.perf_counter()
    for _ in range(reps):
        gp._finalize(th)
    print(json.dumps({"what": "set_theta (final factorisation)", "N": N, "m": m, "ms": 1e3 * (time.perf_counter() - t0) / reps}), flush=True)


if __name__ == "__main__":
    os.environ.setdefault("AOED_FIT_PATH", "blocked")
    cfgs = {"200": (200, 6, 2, [2, 32], 10), "300": (300, 6, 3, [3, 48], 10), "500": (500, 10, 3, [3, 48], 10),
            "1000": (1000, 20, 3, [3, 48], 5), "2000": (2000, 20, 3, [3, 48], 3), "4000": (4000, 32, 3, [3, 24], 2)}
    for k in (sys.argv[1:] or ["200", "300", "500", "1000", "2000", "4000"]):
        run(*cfgs[k])
