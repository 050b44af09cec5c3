import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def isph():
    return importlib.import_module("implicit-sph_b200")


@pytest.fixture(scope="session")
def lattice():
    return importlib.import_module("implicit-sph_b200.lattice")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle as O
    if not os.path.exists(O.PORT_SO):
        O.build(("port",))
    return O
