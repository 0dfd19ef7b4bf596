"""pytest configuration: the ``gpu`` marker, import paths, the xarray stand-in."""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TESTS = os.path.join(ROOT, 'tests')
for p in (ROOT, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)

try:  # real xarray wins when it is installed
    import xarray  # noqa: F401
except ImportError:
    import minixarray
    sys.modules['xarray'] = minixarray


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a real NVIDIA B200 (run with -m gpu on the GPU box)')


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_ok():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
