import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # tests fill the model with dtlr_b200.synth weights right after building it
    config.addinivalue_line("filterwarnings", "ignore:dtlr_b200. no ImageNet ResNet-50 weights found")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
