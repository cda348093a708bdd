import os, sys, cProfile, pstats
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from autooed_b200 import GaussianProcess, ThompsonSampling, HypervolumeImprovement
from autooed_b200.mobo import MOBO
from autooed_b200.problem import SimpleProblem
from autooed_b200.solver import NSGA2
from autooed_b200.utils.sampling import lhs
np.random.seed(0)
prob = SimpleProblem(6, 2)
X = lhs(6, 150); Y = bench._zdt1(X)
sur = GaussianProcess(prob, nu=1)
opt = MOBO(prob, sur, ThompsonSampling(sur), NSGA2(prob, n_gen=200, pop_size=200), HypervolumeImprovement(sur))
opt.optimize(X, Y, None, 10)
pr = cProfile.Profile(); pr.enable()
opt.optimize(X, Y, None, 10)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
