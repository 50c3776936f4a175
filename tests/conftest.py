import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionfinish(session, exitstatus):
    from tests import parity

    parity.dump()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class G:
        def __init__(self):
            self._f = {}

        def __call__(self, name):
            if name not in self._f:
                self._f[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
            return self._f[name]

    return G()
