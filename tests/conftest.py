import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def keys():
    """Reference key fixtures: tests/data/{public,private}_key.bin (reference tests/data/) and the
    network pair (reference src/data/)."""
    from helpers import KeySet

    return KeySet.load()
