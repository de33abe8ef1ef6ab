import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def built_lib():
    """Make sure the in-tree library exists (built by __graft_entry__.build / build.py)."""
    import build
    import torch
    if torch.cuda.is_available() and os.path.exists(build.OUT):
        return build.OUT          # GPU box: use the prebuilt library that travelled with the snapshot
    return build.build(quiet=True)


def rel_l2(a, b):
    import torch
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
