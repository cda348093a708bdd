from autooed_b200.solver.nsga2 import NSGA2
