"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by executing the UNMODIFIED reference
modules from /root/reference (via oracle/ref_loader.py) on seeded inputs.

Run in the authoring container only:   python oracle/make_golden.py
The fixtures are committed; the GPU box never needs /root/reference.

Each bundle holds, for one (problem shape, nu): training data, fitted theta / alpha_ / L_ per objective,
LML + gradient at three thetas (sklearn `log_marginal_likelihood`, `_gpr.py:541-655`), candidate rows and
the outputs of `GaussianProcess.evaluate` (gp.py:141-253 through base.py:68-111) for
  * batched std=False with gradient+hessian,
  * batched std=True (no derivatives),
  * per-row (C=1) std+gradient+hessian stacked — the only form the reference supports,
and of the four acquisitions (identity / ucb / ei / pi) on the per-row stack.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader as rl  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def zdt1(X):
    f1 = X[:, 0]
    g = 1 + 9.0 / (X.shape[1] - 1) * X[:, 1:].sum(1)
    return np.stack([f1, g * (1 - np.sqrt(f1 / g))], 1)  # problem/predefined/zdt.py:28-32


def dtlz2(X, m=3):
    x_, x_m = X[:, : m - 1], X[:, m - 1:]
    g = ((x_m - 0.5) ** 2).sum(1)  # dtlz.py:29-30
    F = []
    for i in range(m):  # dtlz.py:32-44
        f = 1 + g
        f = f * np.prod(np.cos(x_[:, : x_.shape[1] - i] * np.pi / 2.0), axis=1)
        if i > 0:
            f = f * np.sin(x_[:, x_.shape[1] - i] * np.pi / 2.0)
        F.append(f)
    return np.stack(F, 1)


def noisy_sines(X, m, seed):
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((X.shape[1], m))
    return np.sin(X @ W) + 0.1 * (X ** 2) @ np.abs(W) + 0.05 * rng.standard_normal((X.shape[0], m))


def stack_rows(fn, Xs):
    outs = [fn(Xs[i:i + 1]) for i in range(len(Xs))]
    keys = outs[0].keys()
    return {k: (None if outs[0][k] is None else np.concatenate([o[k] for o in outs], 0)) for k in keys}


def make_case(ns, name, X, Y, xl, xu, nu, Xs, seed):
    d, m = X.shape[1], Y.shape[1]
    problem = rl.FakeProblem(d, m, xl, xu)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = rl.fit_reference_gp(ns, problem, X, Y, nu)
    rng = np.random.default_rng(seed + 100)
    bundle = dict(X=X, Y=Y, xl=np.asarray(xl, float), xu=np.asarray(xu, float), nu=np.int64(nu), Xs=Xs)
    Xn, Yn = model.normalization.do(x=X, y=Y)
    bundle["Xn"], bundle["Yn"] = Xn, Yn
    bundle["y_mean"] = model.normalization.y_scaler.mean_
    bundle["y_scale"] = model.normalization.y_scaler.scale_
    thetas, alphas, Ls, lml_star = [], [], [], []
    lml_thetas, lml_vals, lml_grads = [], [], []
    for o, g in enumerate(model.gps):
        thetas.append(g.kernel_.theta.copy())
        alphas.append(g.alpha_.copy())
        Ls.append(g.L_.copy())
        lml_star.append(g.log_marginal_likelihood_value_)
        probes = [g.kernel.theta.copy(), g.kernel_.theta.copy(),
                  rng.uniform(g.kernel_.bounds[:, 0], g.kernel_.bounds[:, 1])]
        pv, pg = [], []
        for th in probes:
            v, gr = g.log_marginal_likelihood(th, eval_gradient=True)
            pv.append(v)
            pg.append(gr)
        lml_thetas.append(np.stack(probes))
        lml_vals.append(np.array(pv))
        lml_grads.append(np.stack(pg))
    bundle.update(theta=np.stack(thetas), alpha=np.stack(alphas), L=np.stack(Ls), lml_star=np.array(lml_star),
                  lml_thetas=np.stack(lml_thetas), lml_vals=np.stack(lml_vals), lml_grads=np.stack(lml_grads))

    def put(prefix, out):
        for k, v in out.items():
            if v is not None:
                bundle[f"{prefix}_{k}"] = v

    # batched, mean only with derivatives (valid for C > 1)
    put("mean", model.evaluate(Xs, std=False, gradient=True, hessian=True))
    # batched mean + std
    for g in model.gps:
        g._K_inv = None
    put("std", model.evaluate(Xs, std=True))
    # per-row full
    put("row", stack_rows(lambda x: model.evaluate(x, std=True, gradient=True, hessian=True), Xs))

    acqs = {"identity": ns.Identity, "ucb": ns.UpperConfidenceBound, "ei": ns.ExpectedImprovement,
            "pi": ns.ProbabilityOfImprovement}
    for aname, cls in acqs.items():
        acq = cls(model)
        acq.fit(X, Y)
        rows = []
        for i in range(len(Xs)):
            F, dF, hF = acq.evaluate(Xs[i:i + 1], gradient=True, hessian=True)
            rows.append((F, dF, hF))
        bundle[f"acq_{aname}_F"] = np.concatenate([r[0] for r in rows])
        bundle[f"acq_{aname}_dF"] = np.concatenate([r[1] for r in rows])
        bundle[f"acq_{aname}_hF"] = np.concatenate([r[2] for r in rows])
        Fb, _, _ = acq.evaluate(Xs)  # batched value-only call (what NSGA-II issues)
        bundle[f"acq_{aname}_Fbatch"] = Fb
    bundle["n_sample"] = np.int64(X.shape[0])
    bundle["y_min"] = Y.min(axis=0)
    path = os.path.join(OUT, f"{name}_nu{nu if nu > 0 else 'rbf'}.npz")
    np.savez_compressed(path, **bundle)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    ns = rl.load()
    os.makedirs(OUT, exist_ok=True)
    for nu in (1, 3, 5, -1):
        rng = np.random.default_rng(10)
        # c1-like: ZDT1 d=6 m=2
        X = rng.random((40, 6))
        Xs = np.vstack([rng.random((6, 6)), X[3:4], np.clip(X[5:6] + 1e-3, 0, 1)])  # incl. a training point (r == 0)
        make_case(ns, "zdt1_d6_n40", X, zdt1(X), np.zeros(6), np.ones(6), nu, Xs, 1)
        # c2-like: DTLZ2 d=10 m=3
        X = rng.random((60, 10))
        Xs = rng.random((6, 10))
        make_case(ns, "dtlz2_d10_n60", X, dtlz2(X, 3), np.zeros(10), np.ones(10), nu, Xs, 2)
        # non-unit bounds + noisy data (well conditioned) d=4 m=2, includes out-of-bounds rows (clipped)
        xl, xu = np.array([-2.0, 0.0, 10.0, -1.0]), np.array([3.0, 5.0, 20.0, 1.0])
        X = xl + rng.random((50, 4)) * (xu - xl)
        Xs = xl + rng.random((6, 4)) * (xu - xl)
        Xs[0, 0] = xu[0] + 1.0
        make_case(ns, "noisy_d4_n50", X, noisy_sines((X - xl) / (xu - xl), 2, 3), xl, xu, nu, Xs, 3)


if __name__ == "__main__":
    main()
