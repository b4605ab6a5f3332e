import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "gpu_sanitize: runs a slice of the GPU suite under compute-sanitizer (slow; -m gpu_sanitize)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def golden_cases(npz, suffix):
    """case names (prefix before the first '_') that have a `<case>_<suffix>` array."""
    return sorted({k[: -len(suffix) - 1] for k in npz.files if k.endswith("_" + suffix)})
