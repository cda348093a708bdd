from autooed_b200.async_strategy.strategies import AsyncStrategy, KrigingBeliever, LocalPenalizer, BelieverPenalizer
