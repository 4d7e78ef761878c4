import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs /root/reference mounted (this container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/src")
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    for it in items:
        if "ref" in it.keywords and not have_ref:
            it.add_marker(pytest.mark.skip(reason="/root/reference not mounted"))
        if "gpu" in it.keywords and not have_gpu:
            it.add_marker(pytest.mark.skip(reason="no CUDA device"))
