import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped automatically when no device is visible (build container)."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def synthetic_model():
    from soccerplayershapepose_b200.model_io import make_synthetic_smpl
    return make_synthetic_smpl(seed=1234)


@pytest.fixture(scope="session")
def wide_model():
    """Harder sparsity statistics: regressors over many mesh parts, skinning rows with 1-4 influences."""
    from soccerplayershapepose_b200.model_io import make_synthetic_smpl
    return make_synthetic_smpl(seed=1234, statistics="wide")


@pytest.fixture(scope="session")
def intree_golden():
    return dict(np.load(os.path.join(GOLDEN_DIR, "intree_golden.npz")))


@pytest.fixture(scope="session")
def smpl_kat():
    return dict(np.load(os.path.join(GOLDEN_DIR, "smpl_kat.npz")))
