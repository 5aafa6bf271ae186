import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


@pytest.fixture(scope='session')
def vms():
    """The product package with a live device context (GPU tests only)."""
    import vaemolsim_b200 as v
    v._abi.ctx()
    return v
