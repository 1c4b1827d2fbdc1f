import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# the cooperative BatchNorm kernel is off by default (see bn_io.cu::bn_fused_min_elems); the tests keep exercising it on
# large inputs (>= 2^22 elements) next to the default two-kernel path on everything smaller
os.environ.setdefault("MOPA_SCN_BN_FUSED_MIN", str(1 << 22))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
