import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pbx_lib():
    from portblas_b200 import build, _lib
    build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def handle(pbx_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    from portblas_b200 import SB_Handle
    h = SB_Handle(0)
    yield h
    h.close()
