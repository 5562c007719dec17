import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, f"{name}.npz"))


@pytest.fixture(scope="session")
def gpu():
    """The CUDA engine.  Fails (does not skip) when the library or the device is missing: the -m gpu
    tests must never pass on a fallback."""
    import fcfc_b200 as F
    F.init()
    return F


@pytest.fixture
def tuned(gpu):
    """Set engine options (fcfc_gpu_set_option) for one test; everything goes back to the defaults afterwards."""
    def set_(name, value=1):
        gpu.set_option(name, value)
    yield set_
    gpu.set_option("defaults", 0)


def have_ref(flavour="dbl_scalar", prog="box"):
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", flavour, f"ref_driver_{prog}"))
