"""Pins the Thompson-sampling oracle (oracle/ts_oracle.py) and the product's RNG-compatible LHS to the reference's own
outputs (`tests/golden/ts_*.npz`, written by oracle/make_golden_ts.py from the unmodified ts.py under numpy seed 11)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import ts_oracle as to

FILES = sorted(f for f in os.listdir(GOLDEN) if f.startswith("ts_"))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.mark.parametrize("name", FILES)
def test_evaluate_on_reference_sample(name):
    g = load(name)
    F, dF, hF = to.ts_evaluate(g["Xs"], g["theta"], g["W"], g["b"], g["sf2"], g["y_mean"], g["y_scale"], True, True)
    assert np.allclose(F, g["F"], rtol=1e-12, atol=1e-12)
    assert np.allclose(dF, g["dF"], rtol=1e-11, atol=1e-11)
    assert np.allclose(hF, g["hF"], rtol=1e-11, atol=1e-10)


@pytest.mark.parametrize("name", FILES)
def test_fit_reproduces_reference_draw(name):
    """Same seed + same GP hyper-parameters => the same spectral points, phases and posterior weight sample."""
    g = load(name)
    np.random.seed(int(g["seed"]))
    theta, W, b, sf2 = to.ts_fit(g["Xn"], g["Yn"], g["gp_theta"], int(g["nu"]))
    assert np.array_equal(np.stack(W), g["W"]) and np.array_equal(np.stack(b), g["b"])
    assert np.allclose(np.stack(theta), g["theta"], rtol=1e-6, atol=1e-8)  # M x M inverse: cond(A) * eps
    assert np.allclose(sf2, g["sf2"], rtol=0, atol=0)


def test_product_lhs_consumes_rng_like_reference():
    from autooed_b200.utils.sampling import lhs
    for n, s in ((3, 7), (1, 100), (6, 100)):
        np.random.seed(5)
        a = to.lhs(n, s)
        state_a = np.random.get_state()[1][:5].copy()
        np.random.seed(5)
        b = lhs(n, s)
        assert np.array_equal(a, b) and np.array_equal(state_a, np.random.get_state()[1][:5])
        assert sorted(np.floor(b[:, 0] * s).astype(int)) == list(range(s))
