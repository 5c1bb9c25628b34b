import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def force_ieee_fp32():
    """torch fp32 references on the GPU must not silently drop to TF32 (cuDNN convs default to TF32)."""
    import torch
    try:
        torch.backends.cudnn.conv.fp32_precision = "ieee"
        torch.backends.cuda.matmul.fp32_precision = "ieee"
    except Exception:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(autouse=True)
def _exact_fp32_references():
    try:
        force_ieee_fp32()
    except Exception:
        pass
    yield


def report(test, **values):
    """Append the MEASURED values behind a tolerance to gpurun_out/parity_report.jsonl (one JSON object per line) when that
    directory exists: every bound in the GPU suite is documented by the number it was derived from (profiles/ keeps a copy)."""
    import json
    out = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out):
        return
    def conv(v):
        try:
            return float(v)
        except Exception:
            return str(v)
    with open(os.path.join(out, "parity_report.jsonl"), "a") as f:
        f.write(json.dumps({"test": test, **{k: conv(v) for k, v in values.items()}}) + "\n")
