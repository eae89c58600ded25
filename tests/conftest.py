import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: long-running CPU oracle test")


@pytest.fixture(scope="session")
def lib():
    from pnpinversion_b200 import _lib

    return _lib.load()


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(0)
    # the fp32 references of the kernel tests are computed by PyTorch on the GPU: keep them true fp32 (cuDNN / cuBLAS
    # would otherwise be free to use TF32, 10 mantissa bits - no better than the fp16 operands under test)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    return torch.device("cuda:0")
