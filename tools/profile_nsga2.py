"""Times one device-resident NSGA-II solve (bench `--mode mobo` configs c1 / c2 shapes); run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split of a generation."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import autooed_b200 as ab
from autooed_b200.problem import SimpleProblem
from autooed_b200.solver import NSGA2
from autooed_b200.utils.sampling import lhs

acq_name, pop, n_gen, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
np.random.seed(0)
prob = SimpleProblem(6, 2)
X = lhs(6, n); Y = bench._zdt1(X)
sur = ab.GaussianProcess(prob, nu=5); sur.fit(X, Y)
acq = ab.get_acquisition(acq_name)(sur); acq.fit(X, Y)
for rep in range(3):
    solver = NSGA2(prob, n_gen=n_gen, pop_size=pop)
    t0 = time.perf_counter()
    solver.solve(X, Y, 10, acq)
    dt = time.perf_counter() - t0
    print(f"{acq_name} pop {pop} gens {n_gen} N {n}: {dt*1e3:.1f} ms, {dt/n_gen*1e6:.0f} us/generation", flush=True)
