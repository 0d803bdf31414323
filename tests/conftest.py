import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return {n: np.load(os.path.join(d, n + ".npz"), allow_pickle=False) for n in ("div", "nms", "search", "tune", "detect", "caffe_layers")}


@pytest.fixture(scope="session")
def golden_blob():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "blob.npz"), allow_pickle=False)
