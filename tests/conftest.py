import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout (build container only)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Both shared libraries are built artefacts (git-ignored); build them if missing/stale.
    nvcc cross-compiles without a GPU, gcc builds the oracle."""
    from gym_roboy_b200 import build as cuda_build
    from oracle import oracle as orc

    cuda_build.build()
    orc.build()


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
