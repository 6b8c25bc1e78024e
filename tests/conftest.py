import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built_libs():
    """Build (or reuse) the in-tree native libraries."""
    from ml_music_style_transfer_b200 import build
    return build.build_all()


@pytest.fixture(scope="session")
def gpu(built_libs):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("test marked gpu but no CUDA device is visible")
    import ml_music_style_transfer_b200 as pkg
    pkg._lib.ops()
    return torch.device("cuda", 0)
