"""Small std+gradient evaluation through the C ABI, used to debug / compare the triangular-product kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autooed_b200 import GaussianProcess  # noqa: E402
from autooed_b200.problem import SimpleProblem  # noqa: E402

N, d, m, C = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (256, 4, 1, 128))]
rng = np.random.default_rng(0)
X = rng.random((N, d))
Y = np.sin(X @ rng.standard_normal((d, m))) + 0.05 * rng.standard_normal((N, m))
thetas = np.stack([np.concatenate([[0.1], rng.uniform(-0.5, 0.5, d), [-4.0]]) for _ in range(m)])
gp = GaussianProcess(SimpleProblem(d, m), nu=5)
gp.fit_fixed(X, Y, thetas)
Xs = rng.random((C, d))
out = gp.evaluate(Xs, std=True, gradient=True)
print("variant", os.environ.get("AOED_TRMM", "tma"), "S", out["S"][:3].ravel(), "dS", out["dS"][0, 0, :3])
np.save(f"gpurun_out/debug_{os.environ.get('AOED_TRMM', 'tma')}.npy", np.concatenate([out["S"].ravel(), out["dS"].ravel()]))
