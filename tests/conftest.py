import os
import sys

import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(REPO, "r-nad_b200")
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
GOLDEN_NAMES = ["cfg1_d2a2c1", "ragged_a3c2", "regular_a3c2d3", "shrinking_a4c2", "c3_a3d2"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=GOLDEN_NAMES)
def golden(request):
    import numpy as np

    data = np.load(os.path.join(GOLDEN_DIR, request.param + ".npz"))
    return request.param, {k: data[k] for k in data.files}
