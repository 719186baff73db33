import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_small():
    return dict(np.load(os.path.join(GOLDEN, "modules_small.npz")))


@pytest.fixture(scope="session")
def golden_pairs():
    return dict(np.load(os.path.join(GOLDEN, "pair_outputs.npz")))


@pytest.fixture(scope="session")
def scans():
    return dict(np.load(os.path.join(GOLDEN, "scans.npz")))


@pytest.fixture(scope="session")
def pretrained_state():
    import torch
    path = os.path.join(GOLDEN, "_big", "rdmnet_state.pt")
    if not os.path.exists(path):
        pytest.skip("pretrained checkpoint copy (tests/golden/_big, git-ignored) not present")
    return torch.load(path, map_location="cpu", weights_only=True)
