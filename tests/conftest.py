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


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a CUDA device AND the built library: skip them (instead of failing in cudaGetDevice) elsewhere"""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = "no CUDA device"
    except Exception as ex:  # noqa: BLE001
        reason = "torch unavailable: %r" % (ex,)
    if reason is None and not os.path.exists(os.path.join(ROOT, "active_tracking_rl_b200", "libtrack2d.so")):
        reason = "libtrack2d.so is not built"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
