import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True))


@pytest.fixture(scope="session")
def load_golden():
    return golden


CONV_CASES = [
    "conv_6_32_M9_K23_B2",
    "conv_64_32_M9_K23",
    "conv_64_64_M8_K16_B2",
    "conv_128_128_M9_K23",
    "conv_32_64_M9_K16_nomask",
    "conv_12_20_M5_K7_trans",
]
