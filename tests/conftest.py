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
    config.addinivalue_line("markers", "live: needs the unmodified reference at /root/reference (build container only)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """Lazy view of one tests/golden/*.npz file with 'case/key' names."""

    def __init__(self, name):
        self._z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    def cases(self):
        return sorted({k.split("/")[0] for k in self._z.files if "/" in k})

    def case(self, name):
        return {k.split("/", 1)[1]: self._z[k] for k in self._z.files if k.startswith(name + "/")}

    def __getitem__(self, k):
        return self._z[k]


@pytest.fixture(scope="session")
def golden_single():
    return Golden("single_env.npz")


@pytest.fixture(scope="session")
def golden_batch():
    return Golden("batch_env.npz")


@pytest.fixture(scope="session")
def golden_gp():
    return Golden("gp.npz")


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-12)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0
