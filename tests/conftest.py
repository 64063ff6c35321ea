import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "refharness")):
    if p not in sys.path:
        sys.path.insert(0, p)

if os.environ.get("TRACK2D_FP32_EMULATION") == "1":  # must precede the first `import torch`
    from active_tracking_rl_b200 import blas
    blas.enable_fp32_emulation(os.environ.get("CUBLAS_EMULATION_STRATEGY", "performant"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


def golden_path(name):
    return os.path.join(GOLDEN, name)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
