import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure). Built by `make -C oracle` / __graft_entry__.build()."""
    from tests import oracle_api

    return oracle_api


@pytest.fixture(scope="session")
def ctx():
    """One device context for the whole GPU session."""
    from kolibrie_b200 import capi

    c = capi.Context(0)
    yield c
    c.close()
