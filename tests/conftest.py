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
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; fails loudly (no CPU fallback) when it or the device is missing."""
    from ipc_b200 import api
    L = api.lib()
    if L.ipc_device_count() <= 0:
        pytest.fail("no CUDA device visible to libipc_b200.so — the gpu-marked tests must run on the GPU box")
    return api
