import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def lib_path():
    """Build (if stale) and return the shared library path.  Building needs nvcc only, not a GPU."""
    from matchnerf_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def ctx(lib_path):
    import torch
    from matchnerf_b200 import capi
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return capi.get_context(torch.device("cuda", 0))
