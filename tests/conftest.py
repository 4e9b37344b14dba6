import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def pre():
    from oracle import oracle, prelude
    oracle.build()
    return prelude


@pytest.fixture(scope="session")
def ort():
    import ort_b200
    return ort_b200


@pytest.fixture(scope="session")
def ctx(ort):
    """A live GPU context; the C ABI fails loudly if no sm_100 GPU is present."""
    c = ort.Context(0)
    yield c
    c.close()
