import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_available() -> bool:
    from alphadia_b200 import _lib

    # a missing / unloadable library is NOT a reason to skip: that must stay a loud failure on the GPU box
    return _lib.load().adb_device_count() >= 1


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not errored) on a machine without a CUDA device; the product code
    itself keeps failing loudly (`_lib.require_device`)."""
    if not any("gpu" in it.keywords for it in items) or _gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle

    oracle.build()
    return oracle
