import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import am_oracle
    am_oracle.build()
    return am_oracle


@pytest.fixture(scope="session")
def native():
    """The built C-ABI library; building it needs nvcc but no GPU."""
    from audio_matcher_b200 import _native
    _native.build_native()
    return _native


@pytest.fixture(scope="session")
def am(native):
    import audio_matcher_b200
    if native.lib().am_device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback to run instead)")
    return audio_matcher_b200
