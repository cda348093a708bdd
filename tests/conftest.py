import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not failed) where no CUDA device is visible, so a plain `pytest` works anywhere."""
    try:
        from autooed_b200 import _lib
        have = _lib.device_count() > 0
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (autooed_b200 has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_files():
    """GP / acquisition fixtures (one per problem shape x nu); pareto_cases.npz is the multi-objective fixture, ts_*.npz
    the Thompson-sampling ones, async_*.npz the penalised-acquisition ones, discovery_*.npz the ParetoDiscovery ones."""
    other = ("ts_", "pareto_", "async_", "discovery_", "pdshell_")
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith(other))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
