import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return load


def rel_err(a, b, floor=1e-30):
    """max-norm relative error per tensor (SURVEY.md 8c parity mode).

    ``floor`` guards tensors that are mathematically zero (e.g. the gradient of a bias that feeds a
    train-mode BatchNorm, or of lin_key.bias under a softmax): there both sides are ~1e-11 noise.
    """
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    if hasattr(b, "detach"):
        b = b.detach().cpu().numpy()
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    den = max(np.abs(b).max(), floor)
    return float(np.abs(a - b).max() / den)


def check_grads(got, want, tol):
    """Every live gradient within ``tol`` relative (max-norm per tensor); tensors whose true value is
    zero are compared against 1e-2 x the largest gradient in the model instead of their own noise."""
    assert set(got) == set(want), (sorted(set(got) ^ set(want)))
    scale = max(float(np.abs(np.asarray(v)).max()) for v in want.values())
    worst = 0.0
    for k in want:
        e = rel_err(got[k], want[k], floor=1e-2 * scale)
        assert e < tol, (k, e)
        worst = max(worst, e)
    return worst
