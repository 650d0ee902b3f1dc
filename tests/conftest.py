import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly when selected without a device; they are deselected
    # with -m "not gpu" on the CPU box, never silently skipped into a fallback.
    pass


@pytest.fixture(scope='session')
def golden():
    g = os.path.join(ROOT, 'tests', 'golden')
    return {k: np.load(os.path.join(g, k + '.npz')) for k in ('weno', 'solver', 'tables')}


def rel_linf(a, b):
    """relative L-infinity error  max|a-b| / max|b|"""
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max())
