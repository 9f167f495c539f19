import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference + numba (build container only)")


@pytest.fixture(scope="session")
def gpu_engine_module():
    from covid19_demography_b200 import engine
    lib = engine.load_library()
    if lib.seir_device_count() < 1:
        pytest.fail("a test marked gpu ran without a CUDA device: the CUDA path has no fallback")
    return engine
