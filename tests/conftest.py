import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_CASES = ['cnn_tiny', 'iqn_tiny', 'cnn_attn', 'iqn_attn', 'iqn_nonorm', 'cnn_selu', 'iqn_selu', 'cnn_elu', 'iqn_tiledz']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN_DIR, f'{name}.pt'), weights_only=False)


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)
