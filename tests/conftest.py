import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _has_gpu() -> bool:
    try:
        from horde_ad_b200 import ffi
        return ffi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a box without a GPU: they are
    # only ever selected with -m gpu, and then they fail loudly if the device
    # or the library is missing (there is no CPU fallback to fall back to).
    pass


@pytest.fixture(scope="session")
def oracle():
    from oracle.concrete import ConcreteBackend
    return ConcreteBackend()


@pytest.fixture(scope="session")
def dev():
    from horde_ad_b200.device import DeviceBackend
    return DeviceBackend()
