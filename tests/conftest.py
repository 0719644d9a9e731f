import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Ref
    if not Ref.available():
        pytest.skip("compiled reference (oracle/_ref) not available")
    return Ref()
