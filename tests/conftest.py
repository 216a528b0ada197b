import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def fixtures():
    from tests import helpers
    return helpers.load_fixtures()


@pytest.fixture(scope="session")
def kats():
    from tests import helpers
    return helpers.load_kats()


@pytest.fixture(scope="session")
def engine():
    """One snp_ctx on cuda:0 for the whole session; fails loudly if the CUDA library is unusable."""
    from snappier_b200.batch import Engine
    e = Engine(0)
    yield e
    e.close()
