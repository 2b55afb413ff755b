import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build the native libraries once (no-op when up to date)."""
    from cpvulkan_b200 import build
    build.build_all()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    from cpvulkan_b200 import capi
    return capi.load_oracle()


@pytest.fixture(scope="session")
def dev(built):
    """The CUDA device through the C ABI. No fallback: a missing library or device is an error, not a skip."""
    from cpvulkan_b200.device import Device
    d = Device(0, stats=True)
    yield d
    d.close()
