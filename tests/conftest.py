import os
import sys
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def kit():
    import cpkit
    cpkit.oracle_lib()
    cpkit.cpsim_lib()
    return cpkit


@pytest.fixture(scope="session")
def hostsim(kit):
    return kit.hostsim_lib()
