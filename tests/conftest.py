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
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def deq():
    import diffeq_b200 as D
    from diffeq_b200 import build as deq_build
    deq_build.build()   # no-op when libdiffeq_b200.so is up to date; nvcc cross-compiles sm_100a without a GPU
    D.lib()             # fails loudly when the extension is missing
    return D
