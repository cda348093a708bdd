#!/usr/bin/env python
"""Timing of the hyper-parameter fit path (north star (d)): batched LML + gradient per device batch,
and full L-BFGS-B fits, next to the CPU oracle (the sklearn formulation restated in numpy) on the same host.
Writes one JSON line per measurement to stdout."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autooed_b200 import GaussianProcess  # noqa: E402
from autooed_b200.problem import SimpleProblem  # noqa: E402
from oracle import gp_oracle as go  # noqa: E402


def data(N, d, m, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.random((N, d))
    Y = np.sin(X @ rng.standard_normal((d, m))) + 0.1 * (X ** 2) @ np.abs(rng.standard_normal((d, m))) \
        + 0.02 * rng.standard_normal((N, m))
    return X, Y


def lml_batch_timing(N, d, m, nu, batches, reps=5, cpu_calls=2):
    X, Y = data(N, d, m)
    Yn = (Y - Y.mean(0)) / Y.std(0)
    gp = GaussianProcess(SimpleProblem(d, m), nu=nu)
    gp._set_train(X, Yn)
    lo, hi = go.theta_bounds(d).T
    rng = np.random.default_rng(1)
    t0 = time.perf_counter()
    for i in range(cpu_calls):
        go.lml_and_grad(rng.uniform(lo, hi) * 0.3, X, Yn[:, i % m], nu)
    cpu_s = (time.perf_counter() - t0) / cpu_calls
    for B in batches:
        obj = np.arange(B) % m
        th = rng.uniform(lo, hi, (B, d + 2)) * 0.3
        gp.log_marginal_likelihood_batch(obj, th, True)  # warm-up
        t0 = time.perf_counter()
        for _ in range(reps):
            gp.log_marginal_likelihood_batch(obj, th, True)
        dt = (time.perf_counter() - t0) / reps
        print(json.dumps({"what": "lml+grad batch", "N": N, "d": d, "nu": nu, "problems": B, "ms_per_batch": 1e3 * dt,
                          "problems_per_s": B / dt, "cpu_oracle_ms_per_problem": 1e3 * cpu_s, "cpu_threads": os.cpu_count(),
                          "speedup_vs_cpu_serial": cpu_s * B / dt, "path": gp.fit_counters()}), flush=True)


def full_fit_timing(N, d, m, nu, n_restarts):
    X, Y = data(N, d, m)
    gp = GaussianProcess(SimpleProblem(d, m), nu=nu, n_restarts=n_restarts, seed=0)
    gp.fit(X, Y)  # warm-up (context, allocations)
    t0 = time.perf_counter()
    gp.fit(X, Y)
    gpu_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc = go.GPOracle(d, m, nu=nu).fit(X, Y)  # one start per objective, as the reference
    cpu_s = time.perf_counter() - t0
    print(json.dumps({"what": "full fit", "N": N, "d": d, "m": m, "nu": nu, "gpu_starts_per_objective": 1 + n_restarts,
                      "gpu_s": gpu_s, "gpu_fit_info": {k: v for k, v in gp.fit_info.items() if k != "nfev"},
                      "cpu_oracle_s_single_start": cpu_s, "cpu_nfev": orc.nfev}), flush=True)


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    lml_batch_timing(200, 6, 2, 5, [1, 2, 16, 148, 592] if not quick else [2, 148])
    lml_batch_timing(500, 10, 3, 5, [3, 48])
    if not quick:
        lml_batch_timing(2000, 20, 3, 5, [3, 48], reps=2, cpu_calls=1)
    full_fit_timing(200, 6, 2, 5, 0)
    full_fit_timing(200, 6, 2, 5, 15)
    full_fit_timing(500, 10, 3, 5, 0)
