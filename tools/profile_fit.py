"""Where a small refit spends its time: wall time per fit, number of LML batches, time inside the C call."""
import os, sys, time, cProfile, pstats
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import autooed_b200 as ab
from autooed_b200.problem import SimpleProblem
from autooed_b200.utils.sampling import lhs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
np.random.seed(0)
prob = SimpleProblem(6, 2)
X = lhs(6, n); Y = bench._zdt1(X)
sur = ab.GaussianProcess(prob, nu=5)
sur.fit(X, Y)
for rep in range(3):
    t0 = time.perf_counter(); sur.fit(X, Y); dt = time.perf_counter() - t0
    print(f"N {n}: fit {dt*1e3:.2f} ms, info {getattr(sur, 'fit_info', None)}", flush=True)
pr = cProfile.Profile(); pr.enable()
for rep in range(5):
    sur.fit(X, Y)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
