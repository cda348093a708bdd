"""TEST INFRASTRUCTURE ONLY — pins the ParetoDiscovery inner loop (local optimisation, KKT duals,
exploration directions, first-order expansion) to the reference: runs the UNMODIFIED functions of
`/root/reference/autooed/mobo/solver/pareto_discovery/pareto_discovery.py:22-341` on a reference-fitted GaussianProcess
with the Identity acquisition under a fixed numpy seed and writes `tests/golden/discovery_*.npz`.

pymoo / pygco are absent, so the module is imported with empty stand-ins for the classes it only subclasses or wires up
(`Algorithm`, `Individual`, the buffer's graph cut ...); none of them is touched by the functions recorded here.  The
evaluation function handed to them follows the pymoo `Problem.evaluate` protocol through
`autooed_b200.surrogate_problem.SurrogateProblem` (host logic) around the REFERENCE acquisition object.

Run in the authoring container only:   python oracle/make_golden_discovery.py
"""
import importlib
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader as rl  # noqa: E402
from oracle.make_golden import zdt1, dtlz2  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def load_reference_module():
    ns = rl.load()

    class _Anything:
        def __init__(self, *a, **k):
            pass

    def stub(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        mod.__path__ = []
        sys.modules[name] = mod

    stub('pymoo'); stub('pymoo.model'); stub('pymoo.model.algorithm', Algorithm=_Anything)
    stub('pymoo.model.duplicate', DefaultDuplicateElimination=_Anything)
    stub('pymoo.model.individual', Individual=_Anything); stub('pymoo.model.initialization', Initialization=_Anything)
    stub('pymoo.optimize', minimize=None); stub('pymoo.factory', get_performance_indicator=None)
    stub('pymoo.performance_indicator'); stub('pymoo.performance_indicator.hv', Hypervolume=_Anything)
    stub('pygco', cut_from_graph=None)
    base = os.path.join(rl.REF_ROOT, "autooed")
    rl._stub('autooed.mobo.solver', os.path.join(base, 'mobo', 'solver'))
    rl._stub('autooed.mobo.solver.pareto_discovery', os.path.join(base, 'mobo', 'solver', 'pareto_discovery'))
    stub('autooed.mobo.solver.base', Solver=_Anything)
    return ns, importlib.import_module('autooed.mobo.solver.pareto_discovery.pareto_discovery')


class _Queue:
    def put(self, item):
        self.item = item


def main():
    from autooed_b200.problem import SimpleProblem
    from autooed_b200.surrogate_problem import SurrogateProblem
    ns, pd = load_reference_module()
    cases = (("zdt1_m2", 5, 6, 2, 40, zdt1), ("dtlz2_m3", 5, 6, 3, 50, lambda X: dtlz2(X, 3)), ("zdt1_nu3_m2", 3, 5, 2, 35, zdt1))
    for name, nu, d, m, N, fn in cases:
        rng = np.random.default_rng(29)
        X = rng.random((N, d))
        Y = fn(X)
        xs = rng.random((7, d))
        xs[0, 0] = 0.0  # an active lower bound from the start
        bounds = np.stack([np.zeros(d), np.ones(d)])
        delta_s, origin_constant, n_grid = 0.3, 1e-2, 6
        origin = np.min(Y, axis=0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = rl.fit_reference_gp(ns, rl.FakeProblem(d, m), X, Y, nu)
            acq = ns.Identity(model)
            acq.fit(X, Y)
            eval_func = SurrogateProblem(SimpleProblem(d, m), acq).evaluate
            # the loop of pareto_discovery.py:321-339, function by function, to record the intermediates ...
            np.random.seed(31)
            ys = eval_func(xs, return_values_of=['F'])
            new_origin = np.minimum(origin, np.min(ys, axis=0))
            if (new_origin != origin).any():
                new_origin -= origin_constant
            fs = ys - new_origin
            x_opts, dirs, samples, patch = [], [], [], []
            for i, (x, y, f) in enumerate(zip(xs, ys, fs)):
                x_opt = pd._local_optimization(x, y, f, eval_func, bounds, None, delta_s)
                directions = pd._get_optimization_directions(x_opt, eval_func, bounds)
                kept = directions.copy()  # :270-283 normalises and rescales the x-part IN PLACE through a view
                x_samples = pd._first_order_approximation(x_opt, directions, bounds, None, n_grid)
                x_opts.append(x_opt); dirs.append(kept); samples.append(x_samples); patch.extend([i] * len(x_samples))
            # ... and the reference's own driver under the same seed must give the same final answer
            np.random.seed(31)
            q = _Queue()
            pd._pareto_discover(xs, eval_func, bounds, None, delta_s, origin, origin_constant, n_grid, q)
            assert np.array_equal(q.item[0], np.vstack(samples)) and q.item[1] == patch and q.item[2] == len(xs)
            assert np.array_equal(q.item[3], new_origin)
            F_opt, dF_opt, hF_opt = eval_func(np.array(x_opts), return_values_of=['F', 'dF']) + (None,)
            hF_opt = np.stack([eval_func(x, return_values_of=['hF']) for x in x_opts])
        rec = dict(X=X, Y=Y, nu=np.int64(nu), gp_theta=np.stack([g.kernel_.theta for g in model.gps]), xs=xs, bounds=bounds,
                   delta_s=delta_s, origin=origin, origin_constant=origin_constant, n_grid=np.int64(n_grid), seed=np.int64(31),
                   ys=ys, new_origin=new_origin, x_opt=np.array(x_opts), F_opt=F_opt, dF_opt=dF_opt, hF_opt=hF_opt,
                   x_samples=np.vstack(samples), patch_ids=np.array(patch))
        for i, dmat in enumerate(dirs):
            rec[f"directions_{i}"] = dmat
        np.savez_compressed(os.path.join(OUT, f"discovery_{name}.npz"), **rec)
        print("wrote discovery_%s.npz" % name, np.vstack(samples).shape, [dm.shape for dm in dirs])


if __name__ == "__main__":
    main()
