import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: skipped (not failed) where there is none."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def synthetic_field(x, y, z, seed=1234):
    """f = sin(3x) cos(2y) cos(z) + 0.1 xi, xi ~ U(-1,1), seed 1234."""
    import numpy as np
    rng = np.random.default_rng(seed)
    return np.asfortranarray(np.sin(3 * x) * np.cos(2 * y) * np.cos(z) + 0.1 * rng.uniform(-1, 1, size=x.shape))


def domain(n, periodic):
    """Node extents the reference decks use: periodic -> [0, 2pi (N-1)/N], bounded -> [0, 1]."""
    import numpy as np
    if periodic:
        return [(0.0, 2 * np.pi * (k - 1) / k) for k in n]
    return [(0.0, 1.0) for _ in n]


def rel_linf(a, b):
    import numpy as np
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle
