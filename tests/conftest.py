import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    """the CPU checker must exist for every test session (compiled in-tree; prebuilt copies travel to the GPU box)"""
    import oracle
    if not os.path.exists(oracle.ORACLE_SO):
        oracle.build(ref=os.path.isdir("/root/reference/src/jams"))
    yield
