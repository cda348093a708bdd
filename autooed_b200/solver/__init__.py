from autooed_b200.solver.nsga2 import NSGA2
from autooed_b200.solver.pareto_discovery import ParetoDiscovery
