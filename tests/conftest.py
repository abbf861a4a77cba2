import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


class Golden:
    """npz fixture written by tests/golden/gen_golden.py (outputs of the reference itself)."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    def __contains__(self, k):
        return k in self.z.files

    def t(self, k):
        a = self.z[k]
        return torch.from_numpy(np.array(a)) if a.shape else torch.tensor(a.item())

    def v(self, k):
        return self.z[k].item()

    def lst(self, k):
        return [self.t(f'{k}.{i}') for i in range(int(self.z[f'{k}.len']))]


@pytest.fixture
def golden():
    return Golden


def load_head_case(name):
    out = Golden(name)
    inp = Golden(f"head_inputs_seed{out.v('seed')}.npz")
    return inp, out
