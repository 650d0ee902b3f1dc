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
    return {k: np.load(os.path.join(g, k + '.npz'))
            for k in ('weno', 'solver', 'solver_sized', 'tables')}


NOISE_FACTOR = 4.


def rel_linf(a, b):
    """relative L-infinity error  max|a-b| / max|b|"""
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(b).max())


def parity_tolerance(golden_solver, name, stated=1e-10):
    """Tolerance of a full-solve parity check: the tolerance BASELINE.json states
    (1e-10 relative L-inf for non-stiff systems, 1e-8 with the stiff Newton solve), or
    4x the reference's own round-off self-noise on that case — the maximum over four
    reference runs with the initial data moved by independent +-1 ulp perturbations,
    stored by tests/golden/make_golden.py — whichever is larger.  The noise comes
    from the reference's wave speeds: spectral radii of forward-difference Jacobians
    (h ~ 1.5e-8)."""
    return max(stated, NOISE_FACTOR * float(golden_solver[name + '__noise']))
