import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(ROOT / "tests" / "golden" / "reproj_golden.npz")


@pytest.fixture(scope="session")
def built_lib():
    """Build (if needed) and load the C-ABI library; never falls back to anything else."""
    from evoworld_b200.build import build_cuda
    from evoworld_b200 import _lib

    build_cuda()
    return _lib.lib()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
