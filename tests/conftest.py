import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure only)."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def recon():
    """A Reconstructor on cuda:0 with the default QM tables uploaded.  GPU tests only."""
    from jxlatte_b200.host import Reconstructor
    r = Reconstructor(0)
    r.generateWeights()
    yield r
    r.close()
