import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'reference: needs the reference tree at /root/reference (authoring container)')


@pytest.fixture(scope='session')
def su_action_mask():
    import torch
    from distar_b200.constants import SELECTED_UNITS_ACTION_MASK
    return torch.tensor(SELECTED_UNITS_ACTION_MASK, dtype=torch.bool)
