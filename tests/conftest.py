import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import ekf_slam_ml_b200
    return ekf_slam_ml_b200


@pytest.fixture(scope="session")
def gpu_pkg(pkg):
    """The product package on a box with a GPU; fails loudly (never skips) when the CUDA path is unusable."""
    assert pkg.device_count() > 0, "no CUDA device visible: -m gpu tests must run on the GPU box"
    return pkg
