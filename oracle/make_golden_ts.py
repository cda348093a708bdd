"""TEST INFRASTRUCTURE ONLY — pins the Thompson-sampling oracle to the reference: runs the UNMODIFIED
`/root/reference/autooed/mobo/acquisition/ts.py` (through oracle/ref_loader.py) on a reference-fitted GaussianProcess
under a fixed numpy seed and writes `tests/golden/ts_*.npz`.

Run in the authoring container only:   python oracle/make_golden_ts.py
"""
import importlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader as rl  # noqa: E402
from oracle.make_golden import zdt1, noisy_sines  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ns = rl.load()
    # ts.py:12 does `from autooed.mobo.surrogate_model import GaussianProcess`: give the stub package that attribute
    sys.modules["autooed.mobo.surrogate_model"].GaussianProcess = ns.GaussianProcess
    TS = importlib.import_module("autooed.mobo.acquisition.ts").ThompsonSampling
    for name, nu, d, m, N in (("zdt1_nu5", 5, 6, 2, 40), ("noisy_nu1", 1, 4, 2, 50), ("noisy_rbf", -1, 4, 2, 50)):
        rng = np.random.default_rng(7)
        X = rng.random((N, d))
        Y = zdt1(X) if name.startswith("zdt1") else noisy_sines(X, m, 3)
        problem = rl.FakeProblem(d, m)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = rl.fit_reference_gp(ns, problem, X, Y, nu)
            acq = TS(model, n_spectral_pts=100)
            np.random.seed(11)
            acq.fit(X, Y)
        Xs = rng.random((9, d))
        F, dF, hF = acq.evaluate(Xs, gradient=True, hessian=True)
        Xn, Yn = model.normalization.do(x=X, y=Y)
        np.savez_compressed(
            os.path.join(OUT, f"ts_{name}.npz"), X=X, Y=Y, Xn=Xn, Yn=Yn, nu=np.int64(nu), Xs=Xs, seed=np.int64(11),
            gp_theta=np.stack([g.kernel_.theta for g in model.gps]), y_mean=model.normalization.y_scaler.mean_,
            y_scale=model.normalization.y_scaler.scale_, theta=np.stack(acq.thetas), W=np.stack(acq.Ws), b=np.stack(acq.bs),
            sf2=np.array(acq.sf2s), F=F, dF=dF, hF=hF)
        print("wrote ts_%s.npz" % name, F.shape, dF.shape, hF.shape)


if __name__ == "__main__":
    main()
