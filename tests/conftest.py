import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def pytest_collection_modifyitems(config, items):
    """A plain ``pytest`` on a box without a CUDA device skips the GPU tests (the engine has
    no CPU fallback, so its fixture would error) instead of failing them."""
    import pytest
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device: genetest_b200 has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
