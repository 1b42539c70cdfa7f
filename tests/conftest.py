import importlib
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    g.build(quiet=True)
    return importlib.import_module("cosig-raytracing_b200")


@pytest.fixture(scope="session")
def abi(pkg):
    return importlib.import_module("cosig-raytracing_b200.abi")


@pytest.fixture(scope="session")
def oracle(pkg):
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py
