import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import Oracle, build
    build(ref=os.path.exists("/root/reference/fieldize.cpp"))
    return Oracle("port")


@pytest.fixture(scope="session")
def ref(port):
    from oracle.oracle import Oracle, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return Oracle("reference")
