import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_sessionstart(session):
    """A fresh checkout has no libhordeb200.so (built artefacts are git-ignored): build it once with the same
    recipe as __graft_entry__.build() when nvcc is on the box (cross-compiles without a GPU).  Without nvcc the
    tests that need the library fail loudly, as they should."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "horde-ad_b200", "libhordeb200.so")
    if os.path.exists(lib) or not (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        return
    env = dict(os.environ)
    env["PATH"] = env.get("PATH", "") + os.pathsep + "/usr/local/cuda/bin"
    subprocess.run(["make", "-C", os.path.join(ROOT, "horde-ad_b200", "csrc"), "-j", str(os.cpu_count() or 4)],
                   check=False, env=env, stdout=subprocess.DEVNULL)


def _has_gpu() -> bool:
    try:
        from horde_ad_b200 import ffi
        return ffi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a box without a GPU.  Selected explicitly with `-m gpu` they run and
    # fail loudly when the device or the library is missing (there is no CPU fallback to fall back to); a plain
    # `pytest tests` on a box without a GPU skips them with the reason spelled out instead of failing 360 tests.
    if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device: run `pytest -m gpu` on a B200 (no CPU fallback exists)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.concrete import ConcreteBackend
    return ConcreteBackend()


@pytest.fixture(scope="session")
def dev():
    from horde_ad_b200.device import DeviceBackend
    return DeviceBackend()
