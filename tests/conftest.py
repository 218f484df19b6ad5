import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
POTENTIALS = os.path.join(ROOT, 'tests', 'potentials')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a CUDA device and the built library: without them they are skipped, not failed."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:      # noqa: BLE001
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason='needs a CUDA device (B200)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def potentials_dir():
    return POTENTIALS
