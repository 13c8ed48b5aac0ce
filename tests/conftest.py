import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def cuda_lib():
    """The C-ABI library on a real device; GPU tests fail loudly when either is missing."""
    import torch
    from voxactb_b200 import _lib
    assert torch.cuda.is_available(), 'gpu-marked test without a CUDA device'
    L = _lib.lib()
    _lib.check(L.vxb_check_device(), 'vxb_check_device')
    return L
