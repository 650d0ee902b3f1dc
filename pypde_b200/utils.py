"""ctypes glue for libpypde.so — mirrors reference pypde/utils.py.

`ADER_ARGTYPES`, `BOUNDARIES`, `nargs`, `c_ptr`, `parse_boundary_types`,
`get_cdll` and `create_solver` keep the reference's names, argument meaning and
error behaviour (pypde/utils.py:9-91).  The library is looked up at
``<package>/build/libpypde.so`` exactly as reference utils.py:69-80 does, and
its absence is a hard error: there is no CPU fallback.
"""
import ctypes
import inspect
import os
import sys
from ctypes import CDLL, POINTER, c_bool, c_double, c_int, c_void_p

from numpy import array

# reference pypde/utils.py:9-16 — unchanged: the three callback slots stay void*
ADER_ARGTYPES = [
    c_void_p, c_void_p, c_void_p, c_bool, c_bool, c_bool,
    POINTER(c_double), c_double,
    POINTER(c_int), c_int,
    POINTER(c_double), c_double,
    POINTER(c_int), c_bool, c_int, c_int, c_int, c_int, c_bool,
    POINTER(c_double), c_int
]

BOUNDARIES = {'transitive': 0, 'periodic': 1}

_LIB = None


def nargs(func):
    return len(inspect.signature(func).parameters)


def c_ptr(arr):
    if arr.dtype == 'int32':
        ptr = POINTER(c_int)
    elif arr.dtype == 'float64':
        ptr = POINTER(c_double)
    else:
        raise TypeError('invalid array type %s' % arr.dtype)
    return arr.ctypes.data_as(ptr)


def parse_boundary_types(boundaryTypes, ndim):
    """reference pypde/utils.py:37-66 (same message, same sys.exit(1))"""
    errMsg = ('boundaryTypes must be a string from {"transitive", "periodic"} '
              'or a list of such strings, of length equal to the number of '
              'dimensions of the domain.')

    if isinstance(boundaryTypes, str):
        try:
            ret = [BOUNDARIES[boundaryTypes]] * ndim
        except KeyError:
            print(errMsg)
            sys.exit(1)
    elif isinstance(boundaryTypes, list):
        if len(boundaryTypes) != ndim:
            print(errMsg)
            sys.exit(1)
        try:
            ret = [BOUNDARIES[b] for b in boundaryTypes]
        except KeyError:
            print(errMsg)
            sys.exit(1)
    else:
        print(errMsg)
        sys.exit(1)

    return array(ret, dtype='int32')


def lib_path():
    loc = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(loc, 'build', 'libpypde.so')


def get_cdll():
    """Loads <package>/build/libpypde.so (reference utils.py:69-80)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                'pypde_b200: %s is missing — build it with '
                '`python -c "import __graft_entry__ as g; g.build()"` or '
                '`make -C pypde_b200/csrc`. There is no CPU fallback.' % path)
        lib = CDLL(path)
        lib.pypde_b200_last_error.restype = ctypes.c_char_p
        lib.pypde_b200_last_error.argtypes = []
        _LIB = lib
    return _LIB


def last_error():
    msg = get_cdll().pypde_b200_last_error()
    return msg.decode() if msg else ''


def check_error(what):
    """The C ABI of the reference returns void; failures are reported here."""
    msg = last_error()
    if msg:
        raise RuntimeError('%s failed: %s' % (what, msg))


def create_solver():
    libpypde = get_cdll()
    solver = libpypde.pde_solver
    solver.argtypes = ADER_ARGTYPES
    solver.restype = None
    return solver
