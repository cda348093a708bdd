"""TEST INFRASTRUCTURE ONLY — pins the penalised-acquisition oracle to the reference: runs the UNMODIFIED
`/root/reference/autooed/mobo/acquisition/penalize.py` (through oracle/ref_loader.py) on a reference-fitted GaussianProcess
under a fixed numpy seed and writes `tests/golden/async_*.npz`.

The reference's asynchronous STRATEGIES cannot be pinned this way: `kb.py` raises NameError (numpy never imported) and
`bp.py:25` reads an unassigned `Y_busy`; `lp.py` only wires `LocalPenalization` up, which is what is recorded here.

Run in the authoring container only:   python oracle/make_golden_async.py
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
    pen = importlib.import_module("autooed.mobo.acquisition.penalize")
    cases = (("zdt1_nu5_identity", 5, 6, 2, 40, ns.Identity), ("noisy_nu1_ucb", 1, 4, 2, 50, ns.UpperConfidenceBound),
             ("noisy_nu3_identity", 3, 5, 3, 45, ns.Identity))
    for name, nu, d, m, N, Acq in cases:
        rng = np.random.default_rng(17)
        X = rng.random((N, d))
        Y = zdt1(X) if name.startswith("zdt1") else noisy_sines(X, m, 3)
        X_busy, Xs = rng.random((3, d)), rng.random((12, d))
        Xs[:2] = X_busy[:2] + 1e-3  # inside the exclusion zones
        problem = rl.FakeProblem(d, m)
        rec = dict(X=X, Y=Y, X_busy=X_busy, Xs=Xs, nu=np.int64(nu), seed=np.int64(23))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = rl.fit_reference_gp(ns, problem, X, Y, nu)
            rec["gp_theta"] = np.stack([g.kernel_.theta for g in model.gps])
            base = Acq(model)
            base.fit(X, Y)
            rec["F_base"] = base.evaluate(Xs)[0]
            rec["busy_mean"], rec["busy_std"] = model.predict(X_busy, std=True)
            np.random.seed(23)
            rec["L_global"] = pen.estimate_lipschitz_constant(model)
            np.random.seed(23)
            rec["L_local"] = pen.estimate_local_lipschitz_constant(model, X_busy)
            for key, cls in (("lp", pen.LocalPenalization), ("llp", pen.LocalLipschitzPenalization),
                             ("hlp", pen.HardLocalPenalization)):
                acq = cls(Acq(model))
                np.random.seed(23)
                acq.fit(X, Y, X_busy)
                rec[key + "_R"], rec[key + "_S"] = acq.R, acq.S
                rec[key + "_F"] = acq.evaluate(Xs)[0]
        np.savez_compressed(os.path.join(OUT, f"async_{name}.npz"), **rec)
        print("wrote async_%s.npz" % name, rec["L_global"], rec["lp_F"][:3].ravel())


if __name__ == "__main__":
    main()
