"""Reference-style user functions (F(Q[, dQ], d) -> ndarray etc., exactly what
pypde.pde_solver accepts) are lowered to CUDA source by symbolic tracing
(pypde_b200/tracing.py).  The generated text is compiled here
with gcc as plain C and compared with the original Python function; where
/root/reference is present the reference's own example systems are traced too."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
from numpy import array, dot, exp, eye, sqrt, where, zeros

from pypde_b200 import cfuncs
from pypde_b200.tracing import TraceError, trace_function


def compile_host(src, kind, tmp_path, tag):
    c = src.replace('extern "C" __device__ ', '')
    path = tmp_path / ('%s.c' % tag)
    path.write_text('#include <math.h>\n#include <stdbool.h>\n' + c)
    so = str(tmp_path / ('%s.so' % tag))
    subprocess.check_call(['gcc', '-O1', '-ffp-contract=off', '-fPIC', '-shared', '-o', so,
                           str(path), '-lm'])
    return getattr(ctypes.CDLL(so), 'user_' + kind)


def call(fn, kind, q, dq, d, nout):
    P = ctypes.POINTER(ctypes.c_double)
    out = np.zeros(nout)
    q = np.ascontiguousarray(q, dtype=float)
    if kind == 'F':
        dq = np.ascontiguousarray(dq, dtype=float)
        fn(out.ctypes.data_as(P), q.ctypes.data_as(P), dq.ctypes.data_as(P), ctypes.c_int(d))
    elif kind == 'B':
        fn(out.ctypes.data_as(P), q.ctypes.data_as(P), ctypes.c_int(d))
    else:
        fn(out.ctypes.data_as(P), q.ctypes.data_as(P))
    return out


# ---- reference-style systems written for this test -------------------------
def F_euler_ref(Q):
    g = 1.4
    r = Q[0]
    E = Q[1] / r
    v = Q[2] / r
    e = E - v**2 / 2
    p = (g - 1) * r * e
    return array([r * v, r * E * v + p * v, r * v**2 + p])


def _pressure(r, E, v):
    return r * 0.4 * (E - dot(v, v) / 2)


def F_ns_ref(Q, dQ, d):
    mu = 1e-2
    ret = zeros(5)
    r = Q[0]
    E = Q[1] / r
    v = Q[2:5] / r
    dv_dx = (dQ[0, 2:5] - dQ[0, 0] * v) / r
    dv = zeros((3, 3))
    dv[0] = dv_dx
    p = _pressure(r, E, v)
    sig = mu * (dv + dv.T - 2 / 3 * (dv[0, 0] + dv[1, 1] + dv[2, 2]) * eye(3))
    vd = v[d]
    ret[0] = r * vd
    ret[1] = r * vd * E + p * vd
    ret[2:5] = r * vd * v
    ret[2 + d] += p
    ret[1] -= dot(sig[d], v)
    ret[2:5] -= sig[d]
    return ret


def B_ref(Q, d):
    ret = zeros((5, 5))
    v = Q[2:5] / Q[0]
    for i in range(2, 5):
        ret[i, i] = v[d]
    ret[2 + d, 2:5] -= v
    return ret


def S_ref(Q):
    ret = zeros(5)
    T = Q[1] / Q[0]
    ret[4] = -Q[4] * 250 * exp(-2 / T) * where(T > 0.25, 1.0, 0.5) + sqrt(abs(Q[0]))
    return ret


@pytest.mark.parametrize('func,kind,ndim,V', [(F_euler_ref, 'F', 1, 3), (F_ns_ref, 'F', 3, 5),
                                              (B_ref, 'B', 3, 5), (S_ref, 'S', 2, 5)])
def test_traced_source_matches_python(func, kind, ndim, V, tmp_path):
    src, second = trace_function(func, kind, ndim, V)
    assert second == (func is F_ns_ref)
    fn = compile_host(src, kind, tmp_path, func.__name__)
    rng = np.random.default_rng(0)
    nargs = func.__code__.co_argcount
    for _ in range(20):
        q = rng.uniform(0.5, 2.0, V)
        dq = rng.standard_normal((ndim, V))
        for d in range(ndim if nargs >= 2 else 1):
            ref = func(q) if nargs == 1 else (func(q, d) if nargs == 2 else func(q, dq, d))
            got = call(fn, kind, q, dq, d, np.size(ref))
            # same operations in the same order; numpy's dot (BLAS) may fuse / reorder
            assert np.allclose(got, np.ravel(ref), rtol=1e-14, atol=1e-15)


# ---- data-dependent branches: rewritten on the source into both-arms-then-select form
def _rate(T):                       # reference tests/reactive_euler/system.py:36-47
    Ti = 0.25
    K0 = 250
    return K0 if T > Ti else 0


def _p_ref(rho):                    # early returns, as reference tests/gpr/misc/mg.py:18-30
    rho0 = 1.2
    if rho > rho0:
        return 3. * (1 / rho0 - 1 / rho) / (1 / rho0 - 0.5 * (1 / rho0 - 1 / rho))**2
    return 3. * (rho - rho0)


def S_branchy(Q):
    ret = zeros(4)
    r = Q[0]
    T = Q[1] / r
    ret[3] = -Q[3] * _rate(T)
    a = 2. * r
    if T > 0.8 and not r > 1.5:     # if / else blocks of assignments, and / not
        a = a + Q[2]
        ret[1] = a * T
    else:
        ret[1] = -a
        ret[2] = 7.
    if Q[2] == Q[2] or T < 0.:      # == on traced values; or
        ret[0] = _p_ref(r)
    if r != r:
        ret[0] = -1.
    return ret


def F_branchy(Q, d):
    ret = zeros(4)
    v = Q[2 + d] / Q[0]             # a branch on the direction stays a Python branch
    if d == 0:
        ret[0] = v
    else:
        ret[0] = -v
    if v > 0.:                      # early return with the rest of the body in the other arm
        ret[1] = Q[1] * v
        return ret
    ret[1] = Q[3] * v
    ret[2] = 1.
    return ret


@pytest.mark.parametrize('func,kind', [(S_branchy, 'S'), (F_branchy, 'F')])
def test_data_dependent_branches_are_lowered(func, kind, tmp_path):
    src, _ = trace_function(func, kind, 2, 4)
    assert ' ? ' in src
    fn = compile_host(src, kind, tmp_path, func.__name__)
    rng = np.random.default_rng(5)
    for _ in range(200):
        q = rng.uniform(0.2, 2.0, 4) * rng.choice([-1., 1.], 4) ** np.array([0, 0, 1, 1])
        for d in range(2 if kind == 'F' else 1):
            ref = func(q, d) if kind == 'F' else func(q)
            got = call(fn, kind, q, np.zeros((2, 4)), d, 4)
            assert np.array_equal(got, np.ravel(ref)), (q, d)


def test_branch_assigned_in_one_arm_only_is_reported():
    def S_bad(Q):
        ret = zeros(3)
        if Q[1] / Q[0] > 0.25:
            k = 250.
        ret[2] = k * Q[2]
        return ret
    with pytest.raises(TraceError, match='only one arm'):
        trace_function(S_bad, 'S', 1, 3)


def test_branch_inside_a_loop_is_reported():
    def S_bad(Q):
        ret = zeros(3)
        for i in range(3):
            if Q[i] > 0.25:
                ret[i] = 1.
        return ret
    with pytest.raises(TraceError, match='where'):
        trace_function(S_bad, 'S', 1, 3)
    with pytest.raises(TypeError, match='cannot lower'):
        cfuncs.generate_cfuncs(None, None, S_bad, 1, 3)


def test_style_detection():
    def F_dev(out, Q, d):                 # device style, 3 parameters, first order
        out[0] = Q[0] * (1.0 + d)
    F, _, _ = cfuncs.generate_cfuncs(F_dev, None, None, 1, 1)
    assert F.style == 'device' and not F.second_order and F.kind == cfuncs.LTOIR
    F, B, S = cfuncs.generate_cfuncs(F_ns_ref, B_ref, S_ref, 3, 5)
    assert F.style == 'reference' and F.second_order and F.kind == cfuncs.CUDA_SOURCE
    assert B.style == 'reference' and S.style == 'reference'
    F, _, _ = cfuncs.generate_cfuncs(F_euler_ref, None, None, 1, 3)
    assert F.style == 'reference' and not F.second_order


def test_traced_functions_link_into_kernels():
    from pypde_b200.utils import get_cdll, last_error
    lib = get_cdll()
    F, B, S = cfuncs.generate_cfuncs(F_ns_ref, B_ref, S_ref, 3, 5)
    n = ctypes.c_size_t()
    rc = lib.pypde_b200_compile(F.pointer, B.pointer, S.pointer, 3, 2, 5, 0, 0, 1,
                                ctypes.byref(n), None, ctypes.c_size_t(0))
    assert rc == 0, last_error()


REF = '/root/reference'


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'pypde')), reason='reference not mounted')
def test_reference_example_systems_trace(tmp_path):
    """The reference's own example systems (pypde/tests/*/system.py), unmodified."""
    import types
    for m in ('matplotlib', 'matplotlib.pyplot', 'mpl_toolkits', 'mpl_toolkits.mplot3d'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['mpl_toolkits.mplot3d'].Axes3D = object
    sys.path.insert(0, REF)
    try:
        from pypde.tests.euler.system import F_euler
        from pypde.tests.gpr.system import B_gpr, F_gpr, S_gpr
        from pypde.tests.navier_stokes.system import F_navier_stokes
        from pypde.tests.reactive_euler.system import F_reactive_euler, S_reactive_euler
    finally:
        sys.path.remove(REF)
    rng = np.random.default_rng(3)

    def gpr_state():
        Q = np.zeros(17)
        Q[0] = rng.uniform(1., 3.)
        Q[2:5] = Q[0] * rng.standard_normal(3) * 0.3
        Q[5:14] = (Q[0]**(1. / 3.) * np.eye(3) + 0.05 * rng.standard_normal((3, 3))).ravel()
        Q[14:17] = Q[0] * rng.standard_normal(3) * 0.01
        Q[1] = Q[0] * 8.0
        return Q

    cases_ = [(F_euler, 'F', 1, 3, lambda: rng.uniform(0.5, 2., 3)),
              (F_reactive_euler, 'F', 2, 6, lambda: rng.uniform(0.5, 2., 6)),
              # (its ignition switch `K0 if T > Ti else 0` lands on both sides over these states)
              (S_reactive_euler, 'S', 2, 6, lambda: rng.uniform(0.5, 2., 6)),
              (F_navier_stokes, 'F', 3, 5, lambda: rng.uniform(0.5, 2., 5)),
              (F_gpr, 'F', 2, 17, gpr_state), (B_gpr, 'B', 2, 17, gpr_state),
              (S_gpr, 'S', 2, 17, gpr_state)]
    for func, kind, ndim, V, state in cases_:
        src, _ = trace_function(func, kind, ndim, V)
        fn = compile_host(src, kind, tmp_path, func.__name__)
        nargs = func.__code__.co_argcount
        for _ in range(5):
            q = state()
            dq = rng.standard_normal((ndim, V))
            for d in range(ndim if nargs >= 2 else 1):
                ref = func(q) if nargs == 1 else (func(q, d) if nargs == 2 else func(q, dq, d))
                got = call(fn, kind, q, dq, d, np.size(ref))
                assert np.allclose(got, np.ravel(ref), rtol=1e-12, atol=1e-13), func.__name__
