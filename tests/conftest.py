import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


if os.environ.get("DTLR_TEST_HALF") == "f16":
    # Run every 16-bit kernel test of the suite against libdtlr_b200_f16.so as well: the tests name torch.bfloat16 explicitly, so this
    # switch aliases it to torch.float16 for the whole session (tolerances written for bf16 are then loose, never tight).
    #     DTLR_TEST_HALF=f16 python -m pytest tests -m gpu
    import torch
    torch.bfloat16 = torch.float16
    torch.Tensor.bfloat16 = torch.Tensor.half


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # tests fill the model with dtlr_b200.synth weights right after building it
    config.addinivalue_line("filterwarnings", "ignore:dtlr_b200. no ImageNet ResNet-50 weights found")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
