import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    try:
        import torch
        torch.backends.cudnn.allow_tf32 = False           # fp32 references must be fp32
        torch.backends.cuda.matmul.allow_tf32 = False
    except Exception:
        pass
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    import json
    import numpy as np
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'reference_golden.npz'))
    with open(os.path.join(ROOT, 'tests', 'golden', 'reference_meta.json')) as f:
        meta = json.load(f)
    return g, meta
