import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_FILES = ["n2_su2_m60_s4.b2seq", "h10_sz_m40_s4.b2seq", "n2_su2_m30_s1.b2seq", "n2_su2_m30_s8.b2seq"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    import b2gpkg
    try:
        return b2gpkg.load().lib().b2g_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def b2g():
    import b2gpkg
    return b2gpkg.load()


@pytest.fixture(scope="session")
def ctx(b2g):
    c = b2g.Context(0)
    yield c
    c.close()
