import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (sm_100a) GPU; run on the GPU box with -m gpu')


@pytest.fixture(scope='session')
def built_lib():
    """Build (or reuse) the in-tree shared library; CPU tests only dlopen it."""
    from vqnerf_release_b200 import build
    return build.build()


@pytest.fixture(scope='session')
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('GPU test selected but no CUDA device is visible (the product has no CPU fallback)')
    return torch.device('cuda:0')
