"""Name -> class maps with the reference's factory names (`autooed/mobo/factory.py:11-22, 25-52, 74-86, 109-115` and
`async_strategy/factory.py:8-29`), so that a maintainer can route 'gp' / 'ei' / 'identity' / 'pi' / 'ucb' / 'hvi' / ... to
the B200 path."""
from autooed_b200.surrogate_model import GaussianProcess
from autooed_b200.acquisition import (Identity, UpperConfidenceBound, ExpectedImprovement, ProbabilityOfImprovement,
                                      ThompsonSampling)
from autooed_b200.acquisition.penalize import LocalPenalization, LocalLipschitzPenalization, HardLocalPenalization
from autooed_b200.async_strategy import KrigingBeliever, LocalPenalizer, BelieverPenalizer
from autooed_b200.selection import HypervolumeImprovement, Uncertainty, Direct, Random
from autooed_b200.solver import NSGA2, ParetoDiscovery

_SURROGATES = {'gp': GaussianProcess}
_ACQUISITIONS = {'ei': ExpectedImprovement, 'identity': Identity, 'pi': ProbabilityOfImprovement,
                 'ts': ThompsonSampling, 'ucb': UpperConfidenceBound}
_SELECTIONS = {'direct': Direct, 'hvi': HypervolumeImprovement, 'random': Random, 'uncertainty': Uncertainty}
_SOLVERS = {'nsga2': NSGA2, 'discovery': ParetoDiscovery}
_ASYNC_ACQUISITIONS = {'lp': LocalPenalization, 'llp': LocalLipschitzPenalization, 'hlp': HardLocalPenalization}
_ASYNC_STRATEGIES = {'kb': KrigingBeliever, 'lp': LocalPenalizer, 'bp': BelieverPenalizer}


def get_surrogate_model(name):
    if name in _SURROGATES:
        return _SURROGATES[name]
    raise Exception(f'Undefined surrogate model {name}')


def get_acquisition(name):
    if name in _ACQUISITIONS:
        return _ACQUISITIONS[name]
    raise Exception(f'Undefined acquisition function {name}')


def get_selection(name):
    if name in _SELECTIONS:
        return _SELECTIONS[name]
    raise Exception(f'Undefined selection {name}')


def get_solver(name):
    if name in _SOLVERS:
        return _SOLVERS[name]
    raise Exception(f'Undefined solver {name}')


def get_async_acquisition(name):
    if name in _ASYNC_ACQUISITIONS:
        return _ASYNC_ACQUISITIONS[name]
    raise Exception(f'Undefined asynchronous acquisition function {name}')


def get_async_strategy(name):
    if name in _ASYNC_STRATEGIES:
        return _ASYNC_STRATEGIES[name]
    raise Exception(f'Undefined asynchronous optimizer {name}')


def _name(name, params, key):
    # the CLI spells module_cfg[...]['surrogate'], GUI configs use ['name'] (SURVEY.md §5 "Config / flag system")
    if name is not None:
        return name
    return params[key] if key in params else params['name']


def init_surrogate_model(name, params, problem):
    return get_surrogate_model(_name(name, params, 'surrogate'))(problem, **params)


def init_acquisition(name, params, surrogate_model):
    return get_acquisition(_name(name, params, 'acquisition'))(surrogate_model, **params)


def init_selection(name, params, surrogate_model):
    return get_selection(_name(name, params, 'selection'))(surrogate_model, **params)


def init_solver(name, params, problem):
    return get_solver(_name(name, params, 'solver'))(problem, **params)


def init_async_acquisition(name, acquisition):
    return get_async_acquisition(name)(acquisition)


def init_async_strategy(params, surrogate_model, acquisition):
    """`params['name']` in {'kb', 'lp', 'bp', 'none'}; the other entries go to the strategy (async_strategy/factory.py:21-29)."""
    assert 'name' in params, 'Name of asynchronous strategy is not specified'
    if params['name'] == 'none':
        return None
    rest = {k: v for k, v in params.items() if k != 'name'}
    return get_async_strategy(params['name'])(surrogate_model, acquisition, **rest)
