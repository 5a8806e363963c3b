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


def _run_two_ranks(mode, out_dir, timeout=600):
    """Two processes of tests/helpers/two_rank_worker.py with a gloo rendezvous on 127.0.0.1; returns what each rank saved."""
    import torch
    import socket
    import subprocess
    import sys
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, 'tests', 'helpers', 'two_rank_worker.py'),
                                       mode, str(out_dir)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    outs = [p.communicate(timeout=timeout)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    return [torch.load(os.path.join(str(out_dir), 'rank%d.pt' % r)) for r in range(2)]


@pytest.fixture
def run_two_ranks():
    return _run_two_ranks
