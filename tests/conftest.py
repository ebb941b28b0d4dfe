import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine; GPU tests must fail (not skip) when the extension cannot be used."""
    from nanomonsv_b200 import Engine
    eng = Engine(0)
    yield eng
    eng.close()
