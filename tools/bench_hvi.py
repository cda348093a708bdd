"""HVI batch selection and Pareto filtering at the c4 candidate count: python tools/bench_hvi.py"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autooed_b200 import _lib  # noqa: E402
from autooed_b200.utils import pareto  # noqa: E402


def main():
    lib = _lib.load()
    for m, C, N, batch in ((2, 1 << 20, 2000, 20), (3, 1 << 20, 2000, 20), (3, 1 << 17, 2000, 20), (2, 1000, 200, 10), (3, 1000, 90, 20)):
        rng = np.random.default_rng(m)
        Y = rng.random((N, m)) * 0.5 + 0.5
        Yc = _lib.as_f64(rng.random((C, m)) ** 0.5)
        t0 = time.perf_counter()
        front = _lib.as_f64(pareto.find_pareto_front(Y))
        t_front = time.perf_counter() - t0
        ref = _lib.as_f64(np.max(np.vstack([Yc, Y]), axis=0))
        idx = np.full(batch, -1, dtype=np.int64)

        def call():
            _lib.check(lib.aoed_hvi_select(0, C, m, _lib.ptr(Yc), len(front), _lib.ptr(front), _lib.ptr(ref), batch, None,
                                           _lib.ptr(idx), None))
        call()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        dt = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        mask = pareto.check_pareto(Yc)
        t_mask = time.perf_counter() - t0
        t0 = time.perf_counter()
        mask = pareto.check_pareto(Yc)
        t_mask = min(t_mask, time.perf_counter() - t0)
        print(json.dumps({"what": "hvi_select", "m": m, "candidates": C, "front": int(len(front)), "batch": batch,
                          "ms_per_call": 1e3 * dt, "ms_per_round": 1e3 * dt / batch,
                          "algorithmic_bytes_per_round": 8 * m * C, "GBps_equiv": 8 * m * C * batch / dt * 1e-9,
                          "h2d_bytes": 8 * m * C, "pareto_mask_ms": 1e3 * t_mask, "pareto_front_size": int(mask.sum()),
                          "find_front_ms_first": 1e3 * t_front}), flush=True)


if __name__ == "__main__":
    main()
