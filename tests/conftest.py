import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure); built on demand from oracle/Makefile."""
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def wen_path():
    return os.path.join(ROOT, "data", "wen.eam.fs")
