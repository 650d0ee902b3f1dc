"""The literal drop-in claim of INTEGRATION.md §1: the reference's OWN Python front end —
pypde/__init__.py, solvers.py, utils.py, byte-compiled unmodified by `make -C oracle
frontend` into oracle/_ref/pypde/*.pyc.bin — runs over this repository's library when

  * pypde_b200/build/libpypde.so is placed where reference utils.py:69-80 looks for it
    (<pypde package>/build/libpypde.so), and
  * pypde/cfuncs.py, the one front-end file that changes, is pypde_b200.cfuncs.

Nothing else is patched: pde_solver is the reference's function, marshalling through the
reference's ADER_ARGTYPES.  Without a GPU the call reaches the library and fails loudly
(CPU test); on a B200 it reproduces the golden Sod solution (GPU test).
"""
import importlib
import os
import shutil
import sys

import numpy as np
import pytest

import cases
from conftest import parity_tolerance, rel_linf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FRONT = os.path.join(ROOT, 'oracle', '_ref', 'pypde')

needs_frontend = pytest.mark.skipif(not os.path.exists(os.path.join(FRONT, 'solvers.pyc.bin')),
                                    reason='oracle/_ref/pypde not built (make -C oracle frontend)')


def F_euler(Q, d):
    """as reference pypde/tests/euler/system.py (reference style: returns an ndarray)"""
    g = 1.4
    r = Q[0]
    E = Q[1] / r
    v = Q[2] / r
    e = E - v**2 / 2
    p = (g - 1) * r * e
    return np.array([r * v, r * E * v + p * v, r * v**2 + p])


@pytest.fixture
def reference_front_end(tmp_path):
    """An installed reference package with the two substitutions of INTEGRATION.md §1."""
    from pypde_b200 import cfuncs
    from pypde_b200.utils import lib_path
    pkg = tmp_path / 'pypde'
    os.makedirs(pkg / 'build')
    for m in ('__init__', 'solvers', 'utils'):       # (*.pyc.bin: see oracle/Makefile)
        shutil.copy(os.path.join(FRONT, m + '.pyc.bin'), pkg / (m + '.pyc'))
    os.symlink(lib_path(), pkg / 'build' / 'libpypde.so')
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'pypde' or
             k.startswith('pypde.')}
    sys.path.insert(0, str(tmp_path))
    sys.modules['pypde.cfuncs'] = cfuncs          # <- the one file that changes
    try:
        pypde = importlib.import_module('pypde')
        assert os.path.dirname(pypde.__file__) == str(pkg)
        yield pypde
    finally:
        sys.path.remove(str(tmp_path))
        for k in [k for k in sys.modules if k == 'pypde' or k.startswith('pypde.')]:
            del sys.modules[k]
        sys.modules.update(saved)


@needs_frontend
def test_reference_front_end_reaches_the_library(reference_front_end):
    """No GPU here: the reference's pde_solver marshals its arguments, binds OUR pde_solver
    through its own ADER_ARGTYPES and calls it; the library reports that it has no CPU path
    (on a GPU box: solves)."""
    from pypde_b200.utils import last_error
    import torch
    pypde = reference_front_end
    Q0 = cases.sod(40)
    out = pypde.pde_solver(Q0, 0.01, [1.], F=F_euler, boundaryTypes='transitive', order=2,
                           ndt=1, stiff=False)
    assert out.shape == (1, 40, 3)
    if torch.cuda.is_available():
        assert last_error() == '' and np.abs(out[0] - cases.sod(40)).max() > 0
    else:
        assert 'CUDA' in last_error() or 'cuda' in last_error()
        assert np.array_equal(out[0], 0 * out[0])       # ret untouched


@pytest.mark.gpu
@needs_frontend
def test_reference_front_end_solves_config1_on_the_gpu(reference_front_end, golden):
    """BASELINE config 1 (Sod, 200 cells, order 2, tf = 0.2) through the reference's own
    pde_solver over libpypde.so, F written as the reference's example writes it."""
    pypde = reference_front_end
    c = cases.solver_cases()['sod_N2']
    Q0 = c['Q0'].copy()
    out = pypde.pde_solver(Q0, c['tf'], c['L'], F=F_euler, boundaryTypes='transitive',
                           order=c['order'], ndt=3, stiff=False)
    from pypde_b200.utils import check_error
    check_error('pde_solver')
    g = golden['solver']
    assert rel_linf(out[-1], g['sod_N2']) < parity_tolerance(g, 'sod_N2')
    assert np.array_equal(Q0, out[-1])                   # Q0 advanced in place
