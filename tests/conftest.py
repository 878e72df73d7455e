import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "unverified: GPU test of code that has not run on a B200 yet; collected LAST so "
                                       "that `-x` never hides the verified tests behind it")
    config.addinivalue_line("markers", "experimental: GPU test of an opt-in kernel path that has never run on a B200; "
                                       "skipped unless GB_EXPERIMENTAL=1")


def pytest_collection_modifyitems(config, items):
    items.sort(key=lambda it: 1 if ("unverified" in it.keywords or "experimental" in it.keywords) else 0)  # stable
    if os.environ.get("GB_EXPERIMENTAL", "0") != "1":
        skip_exp = pytest.mark.skip(reason="opt-in kernel path not yet verified on a B200 (GB_EXPERIMENTAL=1 runs it)")
        for item in items:
            if "experimental" in item.keywords:
                item.add_marker(skip_exp)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
