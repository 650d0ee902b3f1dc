"""Test infrastructure, not product code.

ctypes access to the UNMODIFIED reference built by oracle/Makefile into
oracle/_ref/ (libpypde_ref.so = the reference's own `pde_solver`/`weno_solver`,
libpypde_stages.so = per-stage entry points over the reference classes,
libsystems.so = CPU callbacks compiled from the same C text as the GPU user
functions).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs import this module; nothing here is on the product path.
"""
import ctypes
import os
from ctypes import (CDLL, CFUNCTYPE, POINTER, c_bool, c_double, c_int, c_long,
                    c_void_p)

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, '_ref')

F_TYPE = CFUNCTYPE(None, POINTER(c_double), POINTER(c_double), POINTER(c_double), c_int)
B_TYPE = CFUNCTYPE(None, POINTER(c_double), POINTER(c_double), c_int)
S_TYPE = CFUNCTYPE(None, POINTER(c_double), POINTER(c_double))

# reference pypde/utils.py:9-16
ADER_ARGTYPES = [
    c_void_p, c_void_p, c_void_p, c_bool, c_bool, c_bool,
    POINTER(c_double), c_double,
    POINTER(c_int), c_int,
    POINTER(c_double), c_double,
    POINTER(c_int), c_bool, c_int, c_int, c_int, c_int, c_bool,
    POINTER(c_double), c_int
]
FLUXES = {'rusanov': 0, 'roe': 1, 'osher': 2}
BOUNDARIES = {'transitive': 0, 'periodic': 1}

_libs = {}


def available(name='libpypde_ref.so'):
    return os.path.exists(os.path.join(_REF, name))


def _lib(name):
    if name not in _libs:
        path = os.path.join(_REF, name)
        if not os.path.exists(path):
            raise RuntimeError('%s not built: run `make -C oracle`' % path)
        _libs[name] = CDLL(path)
    return _libs[name]


def dptr(a):
    return a.ctypes.data_as(POINTER(c_double))


def iptr(a):
    return a.ctypes.data_as(POINTER(c_int))


def system_callbacks(system, ndim):
    """Addresses of the gcc-compiled callbacks of oracle/systems.c, as void*."""
    lib = _lib('libsystems.so')

    def addr(kind):
        name = '%s_%dd_%s' % (system, ndim, kind)
        try:
            fn = getattr(lib, name)
        except AttributeError:
            return None
        return ctypes.cast(fn, c_void_p)

    return addr('F'), addr('B'), addr('S')


def _bt(boundaryTypes, ndim):
    if isinstance(boundaryTypes, str):
        boundaryTypes = [boundaryTypes] * ndim
    return np.array([BOUNDARIES[b] for b in boundaryTypes], dtype='int32')


def pde_solver(Q0, tf, L, F=None, B=None, S=None, boundaryTypes='transitive', cfl=0.9,
               order=2, ndt=100, flux='rusanov', stiff=True, nThreads=1, secondOrder=False,
               lib='libpypde_ref.so'):
    """The reference's pde_solver C entry (src/api.h:4-10) with CPU callbacks
    given as void* addresses.  Marshalling as reference solvers.py:177-214.
    Q0 is NOT modified (a copy is advanced)."""
    solver = _lib(lib).pde_solver
    solver.argtypes = ADER_ARGTYPES
    solver.restype = None
    Q0 = np.ascontiguousarray(Q0, dtype='float64')
    nX = np.array(Q0.shape[:-1], dtype='int32')
    ndim = len(nX)
    V = Q0.shape[-1]
    dX = np.array([L[i] / nX[i] for i in range(ndim)], dtype='float64')
    bt = _bt(boundaryTypes, ndim)
    ret = np.zeros(ndt * Q0.size)
    ur = Q0.copy().ravel()
    solver(F, B, S, F is not None, B is not None, S is not None, dptr(ur), tf, iptr(nX), ndim,
           dptr(dX), cfl, iptr(bt), stiff, FLUXES[flux], order, V, ndt, secondOrder, dptr(ret),
           nThreads)
    return ret.reshape((ndt, ) + Q0.shape)


def weno_solver(u, order=2, lib='libpypde_ref.so'):
    """The reference's weno_solver C entry (src/api.h:12-13)."""
    solver = _lib(lib).weno_solver
    solver.argtypes = [POINTER(c_double), POINTER(c_double), POINTER(c_int), c_int, c_int, c_int]
    solver.restype = None
    u = np.ascontiguousarray(u, dtype='float64')
    nX = np.array(u.shape[:-1], dtype='int32')
    ndim = len(nX)
    V = u.shape[-1]
    nXret = nX - 2 * (order - 1)
    ret = np.zeros(int(nXret.prod()) * order**ndim * V)
    solver(dptr(ret), dptr(u.ravel()), iptr(nX), ndim, order, V)
    return ret.reshape(list(nXret) + [order] * ndim + [V])


class Stages:
    """Per-stage calls into the reference classes (oracle/stage_harness.cpp)."""

    def __init__(self, lib='libpypde_stages.so'):
        self.lib = _lib(lib)
        L = self.lib
        L.ref_tables.argtypes = [c_int] + [POINTER(c_double)] * 10
        L.ref_tables.restype = None
        L.ref_boundaries.argtypes = [POINTER(c_double), POINTER(c_double), POINTER(c_int), c_int,
                                     c_int, POINTER(c_int), c_int]
        L.ref_boundaries.restype = None
        L.ref_step.argtypes = [c_void_p, c_void_p, POINTER(c_double), c_long, POINTER(c_double),
                               c_int, c_int, c_int, c_double, c_double, c_int, c_double, c_int]
        L.ref_step.restype = c_double
        L.ref_predictor.argtypes = [POINTER(c_double), c_void_p, c_void_p, c_void_p,
                                    POINTER(c_double), c_long, POINTER(c_double), c_int, c_int,
                                    c_int, c_int, c_double]
        L.ref_predictor.restype = None
        L.ref_fv.argtypes = [POINTER(c_double), c_void_p, c_void_p, c_void_p, POINTER(c_double),
                             c_long, POINTER(c_int), POINTER(c_double), c_int, c_int, c_int, c_int,
                             c_int, c_double]
        L.ref_fv.restype = None
        L.ref_max_abs_eigs.argtypes = [c_void_p, c_void_p, POINTER(c_double), POINTER(c_double),
                                       c_int, c_int, c_int]
        L.ref_max_abs_eigs.restype = c_double
        L.ref_spectral_radius.argtypes = [POINTER(c_double), c_int]
        L.ref_spectral_radius.restype = c_double

    def tables(self, N):
        t = {k: np.zeros(s) for k, s in [('nodes', N), ('wghts', N), ('derv', (N, N)),
                                         ('endv', (2, N)), ('dgmat', (N, N)), ('sig', (N, N)),
                                         ('mL', (N, N)), ('mR', (N, N)), ('mCL', (N, N)),
                                         ('mCR', (N, N))]}
        self.lib.ref_tables(N, *[dptr(t[k]) for k in ['nodes', 'wghts', 'derv', 'endv', 'dgmat',
                                                      'sig', 'mL', 'mR', 'mCL', 'mCR']])
        return t

    def boundaries(self, u, boundaryTypes, N):
        u = np.ascontiguousarray(u, dtype='float64')
        nX = np.array(u.shape[:-1], dtype='int32')
        ndim, V = len(nX), u.shape[-1]
        bt = _bt(boundaryTypes, ndim)
        out = np.zeros(list(nX + 2 * N) + [V])
        self.lib.ref_boundaries(dptr(out), dptr(u), iptr(nX), ndim, V, iptr(bt), N)
        return out

    def step(self, F, B, w, dX, N, cfl, tf, secondOrder, t, count):
        """w: (cells..., N^ndim.., V) any leading shape; returns dt"""
        w = np.ascontiguousarray(w, dtype='float64')
        V = w.shape[-1]
        dX = np.ascontiguousarray(dX, dtype='float64')
        return self.lib.ref_step(F, B, dptr(w), w.size // V, dptr(dX), len(dX), N, V, cfl, tf,
                                 int(secondOrder), t, count)

    def predictor(self, F, B, S, w, dX, N, dt, stiff=False):
        w = np.ascontiguousarray(w, dtype='float64')
        V = w.shape[-1]
        dX = np.ascontiguousarray(dX, dtype='float64')
        qh = np.zeros(w.size * N)
        self.lib.ref_predictor(dptr(qh), F, B, S, dptr(w), w.size // V, dptr(dX), len(dX),
                               int(stiff), N, V, dt)
        return qh

    def fv(self, u, F, B, S, qh, dX, N, dt, flux='rusanov', secondOrder=False):
        u = np.array(u, dtype='float64', order='C', copy=True)
        nX = np.array(u.shape[:-1], dtype='int32')
        V = u.shape[-1]
        dX = np.ascontiguousarray(dX, dtype='float64')
        qh = np.ascontiguousarray(qh, dtype='float64')
        self.lib.ref_fv(dptr(u), F, B, S, dptr(qh), qh.size // V, iptr(nX), dptr(dX), len(nX),
                        FLUXES[flux], N, V, int(secondOrder), dt)
        return u

    def max_abs_eigs(self, F, B, q, dq, d):
        q = np.array(q, dtype='float64')
        dq = np.array(dq, dtype='float64', order='C')
        return self.lib.ref_max_abs_eigs(F, B, dptr(q), dptr(dq), d, q.size, dq.shape[0])

    def spectral_radius(self, M):
        M = np.array(M, dtype='float64', order='C')
        return self.lib.ref_spectral_radius(dptr(M), M.shape[0])
