"""Host-side front end: argument parsing mirrors reference pypde/utils.py and
solvers.py; Python user functions are lowered through numba's CUDA target to
LTO-IR and link into the kernels (no GPU needed for either)."""
import ctypes

import numpy as np
import pytest

from pypde_b200 import cfuncs
from pypde_b200.solvers import FLUXES, _is_second_order
from pypde_b200.utils import (ADER_ARGTYPES, BOUNDARIES, get_cdll, last_error, nargs,
                              parse_boundary_types)


def test_enums_match_reference():
    assert FLUXES == {'rusanov': 0, 'roe': 1, 'osher': 2}        # solvers.py:10
    assert BOUNDARIES == {'transitive': 0, 'periodic': 1}        # utils.py:18
    assert len(ADER_ARGTYPES) == 21                              # utils.py:9-16


def test_parse_boundary_types():
    assert list(parse_boundary_types('periodic', 3)) == [1, 1, 1]
    assert list(parse_boundary_types(['transitive', 'periodic'], 2)) == [0, 1]
    assert parse_boundary_types('transitive', 2).dtype == np.int32
    for bad in ('reflective', ['periodic'], 3):
        with pytest.raises(SystemExit):                          # utils.py:43-64
            parse_boundary_types(bad, 2)


def F_euler1d(out, Q, d):
    g = 1.4
    r = Q[0]
    E = Q[1] / r
    v = Q[2] / r
    e = E - v * v / 2.
    p = (g - 1.) * r * e
    out[0] = r * v
    out[1] = r * E * v + p * v
    out[2] = r * v * v + p


def F_second(out, Q, dQ, d):
    out[0] = Q[0] - 0.1 * dQ[d, 0]


def B_diag(out, Q, d):
    out[0, 0] = 1. + Q[0]


def S_lin(out, Q):
    out[0] = -Q[0]


def test_second_order_by_arity():
    """reference solvers.py:196: secondOrder = (F takes dQ); the lowered function
    carries it (device style: 4 parameters, reference style: 3)."""
    F1, _, _ = cfuncs.generate_cfuncs(F_euler1d, None, None, 1, 3)
    F2, _, _ = cfuncs.generate_cfuncs(F_second, None, None, 1, 1)
    assert not _is_second_order(F1) and _is_second_order(F2)
    assert not _is_second_order(None)


def test_compiled_flux_must_state_its_order():
    """A compiled image shows no arity: neither CudaSource(second_order=) nor the solver's
    secondOrder= given -> refused, never silently first order."""
    src = 'extern "C" __device__ void user_F(double *o, const double *q, const double *dq, int d) {}'
    with pytest.raises(TypeError):
        _is_second_order(cfuncs.CudaSource(src))
    assert _is_second_order(cfuncs.CudaSource(src, second_order=True))
    assert not _is_second_order(cfuncs.CudaSource(src, second_order=False))
    assert _is_second_order(cfuncs.CudaSource(src), secondOrder=True)
    assert not _is_second_order(cfuncs.CudaSource(src, second_order=True), secondOrder=False)
    assert nargs(F_euler1d) == 3


def test_reference_style_function_is_traced():
    def F_ref_style(Q, d):
        return Q * (1.0 + d)
    F = cfuncs._lower(F_ref_style, 'F', 2, 3)
    assert F.kind == cfuncs.CUDA_SOURCE and F.style == 'reference' and b'user_F' in F.image
    with pytest.raises(TypeError, match='output array first'):
        cfuncs.lower_python(F_ref_style, 'F', 1, 3)     # not device style


def test_numba_lowering_links_into_kernels():
    pytest.importorskip('numba')
    lib = get_cdll()
    F = cfuncs.lower_python(F_euler1d, 'F', 1, 3)
    assert F.kind == cfuncs.LTOIR and len(F.image) > 100
    n = ctypes.c_size_t()
    rc = lib.pypde_b200_compile(F.pointer, None, None, 1, 2, 3, 0, 0, 0, ctypes.byref(n), None,
                                ctypes.c_size_t(0))
    assert rc == 0, last_error()
    assert n.value > 10000
    F2, B, S = cfuncs.generate_cfuncs(F_second, B_diag, S_lin, 1, 1)
    rc = lib.pypde_b200_compile(F2.pointer, B.pointer, S.pointer, 1, 2, 1, 0, 0, 1,
                                ctypes.byref(n), None, ctypes.c_size_t(0))
    assert rc == 0, last_error()
