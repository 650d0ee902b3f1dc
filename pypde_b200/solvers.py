"""pde_solver / weno_solver — the reference's Python API (pypde/solvers.py)
over the B200 library.

Argument names, defaults, marshalling and return shapes follow
pypde/solvers.py:13-25,177-214 and :217-244.  Differences, all forced by the
GPU target:

* F, B, S are lowered to device code (pypde_b200/cfuncs.py): reference-style
  Python functions by symbolic tracing, device-style ones through numba's CUDA
  target, or given directly as `CudaSource`; `secondOrder` is still decided by
  F's arity (reference solvers.py:196).
* a failure inside the library raises RuntimeError (the reference's C ABI has
  no error channel; a CUDA library cannot silently continue).
"""
from ctypes import POINTER, c_double, c_int
from multiprocessing import cpu_count

from numpy import array, ascontiguousarray, concatenate, int32, zeros

from pypde_b200.cfuncs import DeviceFunction, generate_cfuncs
from pypde_b200.utils import (c_ptr, check_error, create_solver, get_cdll,
                              nargs, parse_boundary_types)

FLUXES = {'rusanov': 0, 'roe': 1, 'osher': 2}


def _is_second_order(F, secondOrder=None):
    """reference solvers.py:196 decides by F's arity; here the lowered function
    carries the answer (reference style: 3 parameters; device style: 4).  A compiled
    image (CudaSource / DeviceFunction) shows no arity: it must say so itself
    (`second_order=`) or the caller must (`secondOrder=`) — guessing "first order"
    would silently zero dQ in every kernel."""
    if F is None:
        return False
    if secondOrder is not None:
        return bool(secondOrder)
    so = getattr(F, 'second_order', None)
    if so is None:
        raise TypeError('pypde_b200: %s is a compiled device function, so whether F reads dQ '
                        'cannot be seen from its arity: pass second_order=True/False to '
                        'CudaSource / DeviceFunction, or secondOrder=True/False to the solver.'
                        % getattr(F, 'name', 'F'))
    return bool(so)


def pde_solver(Q0,
               tf,
               L,
               F=None,
               B=None,
               S=None,
               boundaryTypes='transitive',
               cfl=0.9,
               order=2,
               ndt=100,
               flux='rusanov',
               stiff=True,
               nThreads=-1,
               wavespeed=None,
               secondOrder=None):
    """Solves dQ/dt + div F(Q, grad Q) + B(Q).grad Q = S(Q) with ADER-WENO on
    the GPU.  Same contract as reference pypde.pde_solver: returns an array of
    shape (ndt,) + Q0.shape; Q0 is advanced in place when it is C-contiguous.

    `wavespeed` (not in the reference; opt-in) is a `CudaSource` defining
    ``extern "C" __device__ double user_L(const double *q, const double *dq, int d)``,
    the analytic max |lambda| of dF_d/dQ + B_d: it replaces the finite-difference
    Jacobian + eigen-solve that the reference — and this library by default — uses
    for the CFL condition and the Rusanov dissipation, so results differ from the
    reference's at the level of its differencing noise (~1e-8 relative in dt).

    `secondOrder` (not in the reference, which reads it off F's arity): needed only
    when F is a compiled `CudaSource` / `DeviceFunction` that does not state
    `second_order` itself; True if F reads dQ.
    """
    nX = array(Q0.shape[:-1], dtype='int32')
    ndim = len(nX)
    V = Q0.shape[-1]
    dX = array([L[i] / nX[i] for i in range(len(L))], dtype='float64')

    boundaryTypes = parse_boundary_types(boundaryTypes, ndim)

    useF = F is not None
    useB = B is not None
    useS = S is not None

    print('compiling functions...')

    _F, _B, _S = generate_cfuncs(F, B, S, ndim, V)

    secondOrder = _is_second_order(_F, secondOrder)

    solver = create_solver()

    ret = zeros(ndt * Q0.size)
    ur = Q0.ravel()

    if nThreads < 1:
        nThreads = cpu_count() - 1

    lib = get_cdll()
    if wavespeed is not None:
        if not isinstance(wavespeed, DeviceFunction):
            raise TypeError('pypde_b200: wavespeed must be a CudaSource / DeviceFunction '
                            'defining user_L')
        if lib.pypde_b200_set_wavespeed(wavespeed.ctypes) != 0:
            check_error('pypde_b200_set_wavespeed')
    try:
        solver(_F.ctypes if useF else None, _B.ctypes if useB else None,
               _S.ctypes if useS else None, useF, useB, useS, c_ptr(ur), tf,
               c_ptr(nX), ndim, c_ptr(dX), cfl, c_ptr(boundaryTypes), stiff,
               FLUXES[flux], order, V, ndt, secondOrder, c_ptr(ret), nThreads)
        check_error('pde_solver')
    finally:
        if wavespeed is not None:
            lib.pypde_b200_set_wavespeed(None)

    return ret.reshape((ndt, ) + Q0.shape)


def _weno_solver_torch(u, order):
    """u: a float64 CUDA tensor; returns a CUDA tensor (nothing crosses PCIe)."""
    import torch
    from ctypes import c_void_p
    u = u.contiguous()
    if u.dtype != torch.float64:
        raise TypeError('pypde_b200.weno_solver: CUDA tensors must be float64')
    nX = array(u.shape[:-1], dtype=int32)
    ndim, V = len(nX), u.shape[-1]
    nXret = nX - 2 * (order - 1)
    ret = torch.empty(tuple(int(x) for x in nXret) + (order, ) * ndim + (V, ),
                      dtype=torch.float64, device=u.device)
    lib = get_cdll()
    lib.pypde_b200_weno_device.argtypes = [c_void_p, c_void_p, POINTER(c_int), c_int, c_int, c_int,
                                           c_void_p]
    with torch.cuda.device(u.device):
        stream = torch.cuda.current_stream().cuda_stream
        if lib.pypde_b200_weno_device(u.data_ptr(), ret.data_ptr(), c_ptr(nX), ndim, order, V,
                                      stream) != 0:
            check_error('weno_solver')
    return ret


def weno_solver(u, order=2):
    """Stand-alone WENO reconstruction (reference solvers.py:217-244).  A float64 CUDA
    torch tensor is reconstructed in place in HBM and a CUDA tensor is returned."""
    if getattr(u, 'is_cuda', False):
        return _weno_solver_torch(u, order)
    u = ascontiguousarray(u, dtype='float64')
    nX = array(u.shape[:-1], dtype=int32)
    ndim = len(nX)
    V = u.shape[-1]

    nXret = nX - 2 * (order - 1)

    libpypde = get_cdll()
    solver = libpypde.weno_solver

    solver.argtypes = [
        POINTER(c_double),
        POINTER(c_double),
        POINTER(c_int),
        c_int,
        c_int,
        c_int,
    ]
    solver.restype = None

    ncellRet = nXret.prod()
    ret = zeros(ncellRet * order**ndim * V)
    ur = u.ravel()

    solver(c_ptr(ret), c_ptr(ur), c_ptr(nX), ndim, order, V)
    check_error('weno_solver')

    return ret.reshape(concatenate([nXret, [order] * ndim, [V]]))
