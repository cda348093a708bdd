"""Where the full c4 fit spends its time: live problems and device time per lock-step round."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from autooed_b200 import GaussianProcess
from autooed_b200.problem import SimpleProblem
X, Y, _ = bench.make_problem(2000, 20, 3)
gp = GaussianProcess(SimpleProblem(20, 3), nu=5, n_restarts=15, seed=3)
log = []
orig = gp.log_marginal_likelihood_batch
def wrapped(obj, th, g=True):
    t0 = time.perf_counter(); out = orig(obj, th, g); log.append((len(obj), time.perf_counter() - t0)); return out
gp.log_marginal_likelihood_batch = wrapped
gp._set_train(X, (Y - Y.mean(0)) / Y.std(0))
gp.log_marginal_likelihood_batch(np.arange(48) % 3, np.zeros((48, 22)), True)  # scratch of the blocked path
log.clear()
if "--second" in sys.argv:  # a second fit of the same object draws other restarts: another round count
    gp.fit(X, Y); log.clear()
t0 = time.perf_counter(); gp.fit(X, Y); wall = time.perf_counter() - t0
n = np.array([l[0] for l in log]); t = np.array([l[1] for l in log])
print(f"wall {wall:.3f}s rounds {len(log)} device-call time {t.sum():.3f}s host {wall - t.sum():.3f}s")
for lo, hi in ((1, 3), (4, 8), (9, 16), (17, 32), (33, 48)):
    m = (n >= lo) & (n <= hi)
    print(f"  live {lo:2d}-{hi:2d}: rounds {m.sum():4d} time {t[m].sum():.3f}s avg {1e3 * t[m].mean() if m.any() else 0:.2f} ms")
print(" ".join(f"{a}:{1e3*b:.1f}" for a, b in zip(n, t)))
