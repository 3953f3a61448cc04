import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build the host stand-in and the oracle (g++); the CUDA library is built by
    __graft_entry__.build() / `python -m eqdyna_b200.build cuda` and only loaded here."""
    from eqdyna_b200 import build
    build.build_host()
    build.build_oracle()
    yield


def _cuda_device_visible():
    return os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly (not skip) when selected with -m gpu on a box without a device: a
    # silent skip there would hide a missing CUDA path.  A plain `pytest tests` on a machine
    # without a device skips them instead of stopping at the first eqd_create error.
    if "gpu" in (config.getoption("-m") or "") or _cuda_device_visible():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (select with -m gpu to make this an error)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
