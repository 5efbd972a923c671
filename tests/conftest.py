import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box: pytest -m gpu)")
    config.addinivalue_line("markers", "ref: needs /root/reference and the oracle/_ref build (build container only)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import port
    port.build()
    return port
