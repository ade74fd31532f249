import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tut3():
    z = np.load(os.path.join(ROOT, "tests", "golden", "tut3_beliefs.npz"))
    return {str(n): (z["pts"][i], z["bw"][i]) for i, n in enumerate(z["names"])}
