"""Host tables (csrc/tables.cpp) and oracle tables against the reference's
(tests/golden/tables.npz, generated from poly/basis.cpp, weno_matrices.cpp,
dg_matrices.cpp) and the known-answer values recorded in SURVEY.md §8c."""
import ctypes

import numpy as np
import pytest

from oracle import ader_weno as O
from pypde_b200.utils import get_cdll


def lib_tables(N):
    lib = get_cdll()
    P = ctypes.POINTER(ctypes.c_double)
    t = {k: np.zeros(s) for k, s in [('nodes', N), ('wghts', N), ('derv', (N, N)),
                                     ('endv', (2, N)), ('dgmat', (N, N)), ('dginv', (N, N)),
                                     ('sig', (N, N)), ('wm', (4, N, N)), ('wminv', (4, N, N))]}
    rc = lib.pypde_b200_tables(N, *[t[k].ctypes.data_as(P) for k in
                                    ['nodes', 'wghts', 'derv', 'endv', 'dgmat', 'dginv', 'sig',
                                     'wm', 'wminv']])
    assert rc == 0
    return t


def test_known_answers_from_survey():
    t = lib_tables(2)
    assert np.allclose(t['nodes'], [0.21132486540518713, 0.78867513459481287], rtol=0, atol=1e-16)
    assert np.allclose(t['wghts'], [0.5, 0.5], rtol=0, atol=1e-16)
    assert np.allclose(np.abs(t['derv']), 1.7320508075688774, rtol=1e-15)
    assert np.allclose(t['dgmat'], [[1, 0.36602540378443837], [-1.366025403784439, 1.0000000000000002]],
                       rtol=0, atol=1e-15)
    t = lib_tables(3)
    assert np.allclose(t['nodes'], [0.1127016653792583, 0.5, 0.8872983346207417], rtol=0, atol=1e-16)
    assert np.allclose(t['derv'][0], [-3.8729833462074175, 5.1639777949432242, -1.2909944487358058],
                       rtol=1e-15)
    assert np.allclose(t['endv'][0], [1.4788305577012362, -0.66666666666666652, 0.18783610896543049],
                       rtol=0, atol=2e-15)
    assert np.allclose(t['sig'][0], [49.814814814814802, -96.296296296296276, 46.481481481481467],
                       rtol=1e-15)
    assert abs(t['sig'][1, 1] - 192.59259259259255) < 1e-12


@pytest.mark.parametrize('N', [2, 3, 4])
def test_against_reference_tables(golden, N):
    t = lib_tables(N)
    o = O.tables(N)
    g = golden['tables']
    for k in ['nodes', 'wghts', 'derv', 'endv', 'dgmat', 'sig']:
        ref = g['N%d_%s' % (N, k)]
        scale = max(1., np.abs(ref).max())
        assert np.abs(t[k] - ref).max() / scale < 4e-15, k
        assert np.abs(getattr(o, k) - ref).max() / scale < 4e-15, k
    for i, k in enumerate(['mL', 'mR', 'mCL', 'mCR']):
        ref = g['N%d_%s' % (N, k)]
        assert np.abs(t['wm'][i] - ref).max() / np.abs(ref).max() < 4e-15, k
        assert np.abs(o.wm[i] - ref).max() / np.abs(ref).max() < 4e-15, k
    # the inverses the kernels use
    for i in range(4):
        assert np.abs(t['wminv'][i] @ t['wm'][i] - np.eye(N)).max() < 1e-12
    assert np.abs(t['dginv'] @ t['dgmat'] - np.eye(N)).max() < 1e-14


@pytest.mark.parametrize('N', [1, 2, 3, 4, 5, 6])
def test_quadrature_exactness(N):
    t = lib_tables(N)
    # Gauss rule on [0,1]: exact for degree < 2N
    for k in range(2 * N):
        assert abs(np.dot(t['wghts'], t['nodes']**k) - 1. / (k + 1)) < 1e-15
    # Lagrange basis: derivative matrix rows sum to 0, end values sum to 1
    assert np.abs(t['derv'].sum(axis=1)).max() < 1e-12
    assert np.abs(t['endv'].sum(axis=1) - 1).max() < 1e-13
