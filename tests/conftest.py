import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def weights():
    from oracle import seam_oracle as so
    return so.random_weights(seed=0)


@pytest.fixture(scope="session")
def engine(weights):
    """One SeamEngine on cuda:0 with the golden weights loaded (GPU tests only)."""
    import torch
    import seam_match_rcnn_b200 as pkg
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    e = pkg.SeamEngine("cuda:0")
    e.load_weights({k: v.cuda() for k, v in weights.items()})
    return e
