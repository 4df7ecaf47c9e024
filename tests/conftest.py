import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


@pytest.fixture(scope='session')
def golden():
    import numpy as np
    d = os.path.join(ROOT, 'tests', 'golden')
    return {name: np.load(os.path.join(d, name + '.npz')) for name in ('corr', 'raft', 'warp', 'mask', 'greedy')}


@pytest.fixture(scope='session')
def lib():
    """libsdof_b200.so, built on demand (nvcc cross-compiles without a GPU)."""
    from sd_animation_optical_flow_b200 import _capi, build
    build.build()
    return _capi.load()


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('a test marked gpu ran without a CUDA device')
    from sd_animation_optical_flow_b200 import build
    build.build()
    return torch.device('cuda', 0)
