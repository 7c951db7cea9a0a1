import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def cuda_ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from custom_d_fine_b200.cuda_ops import CudaOps
    return CudaOps()


@pytest.fixture(scope="session")
def oracle_ops():
    from oracle.torch_ops import OracleOps
    return OracleOps()
