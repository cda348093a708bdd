"""autooed_b200 — B200-native (sm_100a) GP-surrogate hot path of AutoOED behind the autooed.mobo interfaces.

    from autooed_b200 import GaussianProcess, ExpectedImprovement, HypervolumeImprovement

All numerics run in libaoed_b200.so (hand-written CUDA, C ABI in include/aoed_b200.h); there is no CPU fallback.
"""
from autooed_b200.surrogate_model import GaussianProcess
from autooed_b200.acquisition import (Identity, UpperConfidenceBound, ExpectedImprovement, ProbabilityOfImprovement,
                                      ThompsonSampling)
from autooed_b200.selection import HypervolumeImprovement, Uncertainty, Direct, Random
from autooed_b200.solver import NSGA2
from autooed_b200.acquisition.penalize import LocalPenalization, LocalLipschitzPenalization, HardLocalPenalization
from autooed_b200.async_strategy import KrigingBeliever, LocalPenalizer, BelieverPenalizer
from autooed_b200.factory import (get_surrogate_model, get_acquisition, get_selection, get_solver, init_surrogate_model,
                                  init_acquisition, init_selection, init_solver, get_async_acquisition,
                                  init_async_acquisition, get_async_strategy, init_async_strategy)

__version__ = "0.1.0"
