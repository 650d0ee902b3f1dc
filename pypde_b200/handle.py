"""Handle API (include/pypde_b200.h, part 2) for state resident in HBM.

`Solver` drives the same C++ host driver that `pde_solver` uses, one time step
at a time, on a caller-chosen CUDA stream and, optionally, in place on a
caller-owned device buffer (e.g. a torch tensor).  Used by bench.py and by the
stage-wise parity tests.
"""
import ctypes
from ctypes import POINTER, byref, c_double, c_int, c_longlong, c_size_t, c_void_p

import numpy as np

from pypde_b200.cfuncs import generate_cfuncs
from pypde_b200.solvers import FLUXES, _is_second_order
from pypde_b200.utils import get_cdll, last_error, parse_boundary_types

_configured = False


def _lib():
    global _configured
    lib = get_cdll()
    if not _configured:
        P = c_void_p
        lib.pypde_b200_create.argtypes = [POINTER(c_void_p), P, P, P, POINTER(c_int), c_int,
                                          POINTER(c_double), c_double, POINTER(c_int), c_int,
                                          c_int, c_int, c_int, c_int]
        lib.pypde_b200_destroy.argtypes = [c_void_p]
        lib.pypde_b200_set_stream.argtypes = [c_void_p, c_void_p]
        lib.pypde_b200_set_state.argtypes = [c_void_p, POINTER(c_double)]
        lib.pypde_b200_get_state.argtypes = [c_void_p, POINTER(c_double)]
        lib.pypde_b200_bind_state.argtypes = [c_void_p, c_void_p]
        lib.pypde_b200_begin.argtypes = [c_void_p, c_double]
        lib.pypde_b200_step_async.argtypes = [c_void_p]
        lib.pypde_b200_sync.argtypes = [c_void_p, POINTER(c_double), POINTER(c_double),
                                        POINTER(c_int)]
        lib.pypde_b200_launch_count.argtypes = [c_void_p]
        lib.pypde_b200_launch_count.restype = c_longlong
        lib.pypde_b200_read_stage.argtypes = [c_void_p, c_int, POINTER(c_double), c_size_t,
                                              POINTER(c_size_t)]
        lib.pypde_b200_set_profiling.argtypes = [c_void_p, c_int]
        lib.pypde_b200_kernel_times.argtypes = [c_void_p, ctypes.c_char_p, c_size_t]
        lib.pypde_b200_fp64_peak.argtypes = [c_void_p, POINTER(c_double)]
        lib.pypde_b200_comm_unique_id.argtypes = [c_void_p]
        lib.pypde_b200_comm_init.argtypes = [c_int, c_int, c_void_p]
        _configured = True
    return lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError('%s failed: %s' % (what, last_error()))


STAGES = {'ub': 0, 'w': 1, 'traces': 2, 'centers': 3, 'flux0': 4, 'flux1': 5, 'flux2': 6, 'wavespeeds': 7,
          'stiff_stats': 8}


class Solver:
    """One slab of the domain on one GPU.

    shape = (nX_0, ..., nX_{ndim-1}, V); L = domain lengths (used as dX = L/nX,
    as reference solvers.py:180); F/B/S as for `pde_solver`.
    """

    def __init__(self, shape, L, F=None, B=None, S=None, boundaryTypes='transitive', cfl=0.9,
                 order=2, flux='rusanov', stiff=False, dX=None, secondOrder=None):
        self.lib = _lib()
        self.shape = tuple(int(s) for s in shape)
        nX = np.array(self.shape[:-1], dtype='int32')
        self.ndim = len(nX)
        self.V = self.shape[-1]
        self.N = order
        if dX is None:
            dX = [L[i] / nX[i] for i in range(self.ndim)]
        self.dX = np.array(dX, dtype='float64')
        bt = parse_boundary_types(boundaryTypes, self.ndim)
        self._fns = generate_cfuncs(F, B, S, self.ndim, self.V)
        ptrs = [f.ctypes if f is not None else None for f in self._fns]
        self.h = c_void_p()
        _check(self.lib.pypde_b200_create(byref(self.h), ptrs[0], ptrs[1], ptrs[2],
                                          nX.ctypes.data_as(POINTER(c_int)), self.ndim,
                                          self.dX.ctypes.data_as(POINTER(c_double)), cfl,
                                          bt.ctypes.data_as(POINTER(c_int)), int(stiff),
                                          FLUXES[flux], order, self.V,
                                          int(_is_second_order(self._fns[0], secondOrder))),
               'pypde_b200_create')
        self.ncell = int(nX.prod())
        self._bound = None

    def close(self):
        if self.h:
            self.lib.pypde_b200_destroy(self.h)
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        _check(self.lib.pypde_b200_set_stream(self.h, c_void_p(stream)), 'set_stream')

    def set_state(self, u):
        u = np.ascontiguousarray(u, dtype='float64')
        assert u.size == self.ncell * self.V
        _check(self.lib.pypde_b200_set_state(self.h, u.ctypes.data_as(POINTER(c_double))),
               'set_state')

    def get_state(self):
        u = np.zeros(self.shape)
        _check(self.lib.pypde_b200_get_state(self.h, u.ctypes.data_as(POINTER(c_double))),
               'get_state')
        return u

    def bind_tensor(self, tensor):
        """Advance a CUDA float64 torch tensor of shape `shape` in place."""
        assert tensor.is_cuda and tensor.is_contiguous() and tensor.numel() == self.ncell * self.V
        self._bound = tensor
        _check(self.lib.pypde_b200_bind_state(self.h, c_void_p(tensor.data_ptr())), 'bind_state')

    def begin(self, tf):
        _check(self.lib.pypde_b200_begin(self.h, tf), 'begin')

    def step_async(self):
        _check(self.lib.pypde_b200_step_async(self.h), 'step_async')

    def sync(self):
        t, dt, nan = c_double(), c_double(), c_int()
        _check(self.lib.pypde_b200_sync(self.h, byref(t), byref(dt), byref(nan)), 'sync')
        return t.value, dt.value, bool(nan.value)

    def step(self):
        self.step_async()
        return self.sync()

    def set_profiling(self, on):
        _check(self.lib.pypde_b200_set_profiling(self.h, int(on)), 'set_profiling')

    def kernel_times(self):
        """{kernel name: (total device ms, launches)} since the last call."""
        buf = ctypes.create_string_buffer(4096)
        _check(self.lib.pypde_b200_kernel_times(self.h, buf, 4096), 'kernel_times')
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split()
            out[name] = (float(ms), int(n))
        return out

    def fp64_peak_tflops(self):
        v = c_double()
        _check(self.lib.pypde_b200_fp64_peak(self.h, byref(v)), 'fp64_peak')
        return v.value

    @property
    def launches(self):
        return int(self.lib.pypde_b200_launch_count(self.h))

    def stiff_stats(self):
        """k_dg_stiff's iteration counters since the solver was created (needs
        PYPDE_B200_STIFF_STATS=1 when the solver is built): Newton iterations, evaluations
        of the predictor's objective, inner Krylov steps, and inner steps beyond the
        shared-memory resident basis."""
        raw = self.read_stage('stiff_stats').view(np.uint64)
        return dict(zip(('newton', 'obj', 'inner', 'deep'), (int(x) for x in raw)))

    def read_stage(self, name):
        n = c_size_t()
        _check(self.lib.pypde_b200_read_stage(self.h, STAGES[name], None, 0, byref(n)),
               'read_stage')
        out = np.zeros(n.value)
        _check(self.lib.pypde_b200_read_stage(self.h, STAGES[name],
                                              out.ctypes.data_as(POINTER(c_double)), n.value,
                                              byref(n)), 'read_stage')
        return out


def comm_init_from_torch():
    """Builds the NCCL slab communicator from an initialised torch.distributed
    process group (torch is plumbing: it only ferries the 128-byte unique id)."""
    import torch
    import torch.distributed as dist
    lib = _lib()
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        _check(lib.pypde_b200_comm_unique_id(buf), 'comm_unique_id')
    t = torch.tensor(list(buf.raw), dtype=torch.uint8)
    if dist.get_backend() == 'nccl':
        t = t.cuda()
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    buf2 = ctypes.create_string_buffer(raw, 128)
    _check(lib.pypde_b200_comm_init(rank, world, buf2), 'comm_init')
    return rank, world
