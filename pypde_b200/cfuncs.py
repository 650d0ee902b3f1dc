"""User-function lowering — the one front-end file that differs from the
reference (pypde/cfuncs.py:30-75).

The reference njit-compiles F/B/S and wraps them in numba `@cfunc` CPU
callbacks.  Here `generate_cfuncs` returns, with the same call signature and
the same `(_F, _B, _S)` result shape, *device-function descriptors*
(`pypde_b200_devfn`, include/pypde_b200.h) whose `.ctypes` attribute is passed
through the unchanged `c_void_p` slots of `ADER_ARGTYPES`.  Each descriptor
holds one of

* LTO-IR produced by numba's CUDA target from a Python function
  (`numba.cuda.compile(..., device=True, abi='c', output='ltoir')`),
* CUDA C++ source (compiled by NVRTC inside the library), or
* PTX,

and the library links it into the hand-written kernels with nvJitLink.

Python functions come in two styles:

* reference style — exactly what the reference accepts (pypde/solvers.py:30-55):
  ``F(Q)``, ``F(Q, d)``, ``F(Q, dQ, d)``, ``B(Q)``, ``B(Q, d)``, ``S(Q)`` returning
  arrays.  numba's CUDA target cannot allocate or return arrays (SURVEY.md
  §7.3-H2), so these are lowered by symbolic tracing to CUDA source
  (pypde_b200/tracing.py);
* device style — scalar code writing into a leading `out` argument, lowered
  through numba's CUDA target to LTO-IR:
  ``F(out, Q, d)`` / ``F(out, Q, dQ, d)``, ``B(out, Q, d)``, ``S(out, Q)``.

A three-parameter F is ambiguous (``F(Q, dQ, d)`` or ``F(out, Q, d)``): it is
traced as reference style first; if it returns nothing it is device style.
"""
import ctypes
from ctypes import POINTER, Structure, c_char_p, c_int, c_size_t, c_void_p

LTOIR = 0
PTX = 1
CUDA_SOURCE = 2


class _DevFn(Structure):
    _fields_ = [('image', c_void_p), ('bytes', c_size_t), ('kind', c_int),
                ('name', c_char_p)]


class DeviceFunction:
    """A device function image + the C descriptor that points at it."""

    def __init__(self, image, kind, name, second_order=None):
        # second_order (F only): True if the function reads dQ — the reference decides this
        # by F's arity (solvers.py:196), which a compiled image does not show.  None =
        # not stated: pde_solver then needs its secondOrder= argument.
        self.second_order = second_order
        if isinstance(image, str):
            image = image.encode()
        if kind in (PTX, CUDA_SOURCE) and not image.endswith(b'\0'):
            image = image + b'\0'
        self.image = image
        self.kind = kind
        self.name = name
        self._buf = ctypes.create_string_buffer(image, len(image))
        self._name = ctypes.create_string_buffer(name.encode())
        self._desc = _DevFn(ctypes.cast(self._buf, c_void_p), len(image), kind,
                            ctypes.cast(self._name, c_char_p))

    @property
    def ctypes(self):
        """What solvers.py passes in the callback slot (reference: `_F.ctypes`)."""
        return ctypes.cast(ctypes.pointer(self._desc), c_void_p)

    @property
    def pointer(self):
        return ctypes.pointer(self._desc)


class CudaSource(DeviceFunction):
    """CUDA C++ text defining `extern "C" __device__ void user_F/B/S(...)`."""

    def __init__(self, source, name='user_function', second_order=None):
        super().__init__(source, CUDA_SOURCE, name, second_order)


class _offline_device:
    """numba asks the current device for its compute capability when a device
    function calls another one; LTO-IR is architecture-neutral, so on a machine
    without a driver (build box, CPU tests) a stand-in answers (9, 0) — the
    highest target numba 0.65 knows; nvJitLink generates the sm_100a code."""

    class _Dev:
        compute_capability = (9, 0)

    def __enter__(self):
        from numba.cuda import dispatcher
        self._mod = dispatcher
        self._orig = dispatcher.get_current_device
        orig = self._orig

        def get_device():
            try:
                return orig()
            except Exception:
                return _offline_device._Dev()

        dispatcher.get_current_device = get_device

    def __exit__(self, *exc):
        self._mod.get_current_device = self._orig
        return False


def _device_style_arity(kind):
    return {'F': (3, 4), 'B': (3, ), 'S': (2, )}[kind]


def lower_python(func, kind, ndim, V):
    """Python device-style function -> LTO-IR with a C ABI named user_<kind>."""
    from numba import cuda, types
    from pypde_b200.utils import nargs

    n = nargs(func)
    if n not in _device_style_arity(kind):
        raise TypeError(
            'pypde_b200: %s has %d parameters; device-style GPU user functions take the '
            'output array first — F(out, Q, d) / F(out, Q, dQ, d), B(out, Q, d), '
            'S(out, Q) — and use scalar arithmetic only.' %
            (getattr(func, '__name__', kind), n))

    dev = cuda.jit(device=True, inline=True)(func)
    dptr = types.CPointer(types.float64)

    if kind == 'F':
        if n == 3:

            def wrapper(out, q, dq, d):
                Q = cuda.local.array(V, types.float64)
                R = cuda.local.array(V, types.float64)
                for i in range(V):
                    Q[i] = q[i]
                dev(R, Q, d)
                for i in range(V):
                    out[i] = R[i]
        else:

            def wrapper(out, q, dq, d):
                Q = cuda.local.array(V, types.float64)
                DQ = cuda.local.array((ndim, V), types.float64)
                R = cuda.local.array(V, types.float64)
                for i in range(V):
                    Q[i] = q[i]
                for k in range(ndim):
                    for i in range(V):
                        DQ[k, i] = dq[k * V + i]
                dev(R, Q, DQ, d)
                for i in range(V):
                    out[i] = R[i]

        sig = types.void(dptr, dptr, dptr, types.int32)
    elif kind == 'B':

        def wrapper(out, q, d):
            Q = cuda.local.array(V, types.float64)
            R = cuda.local.array((V, V), types.float64)
            for i in range(V):
                Q[i] = q[i]
            for i in range(V):
                for j in range(V):
                    R[i, j] = 0.
            dev(R, Q, d)
            for i in range(V):
                for j in range(V):
                    out[i * V + j] = R[i, j]

        sig = types.void(dptr, dptr, types.int32)
    else:

        def wrapper(out, q):
            Q = cuda.local.array(V, types.float64)
            R = cuda.local.array(V, types.float64)
            for i in range(V):
                Q[i] = q[i]
            for i in range(V):
                R[i] = 0.
            dev(R, Q)
            for i in range(V):
                out[i] = R[i]

        sig = types.void(dptr, dptr)

    with _offline_device():
        ltoir, _ = cuda.compile(wrapper, sig, device=True, abi='c',
                                abi_info={'abi_name': 'user_' + kind},
                                output='ltoir', cc=(9, 0))
    return DeviceFunction(bytes(ltoir), LTOIR,
                          getattr(func, '__name__', 'user_' + kind))


REFERENCE_ARITY = {'F': (1, 2, 3), 'B': (1, 2), 'S': (1, )}


def lower_reference_style(func, kind, ndim, V):
    """Reference-style function -> CUDA source by symbolic tracing."""
    from pypde_b200.tracing import trace_function
    src, second_order = trace_function(func, kind, ndim, V)
    fn = CudaSource(src, getattr(func, '__name__', 'user_' + kind))
    fn.second_order = second_order
    fn.style = 'reference'
    return fn


def _lower(func, kind, ndim, V):
    if func is None:
        return None
    if isinstance(func, DeviceFunction):
        return func
    from pypde_b200.utils import nargs
    n = nargs(func)
    trace_err = None
    if n in REFERENCE_ARITY[kind]:
        try:
            return lower_reference_style(func, kind, ndim, V)
        except Exception as err:
            trace_err = err
            if n not in _device_style_arity(kind):
                raise TypeError('pypde_b200: cannot lower %s for the GPU: %s' %
                                (getattr(func, '__name__', kind), err)) from err
    try:
        fn = lower_python(func, kind, ndim, V)
    except Exception as err:
        if trace_err is None:
            raise
        raise TypeError('pypde_b200: cannot lower %s for the GPU.\n  as a reference-style '
                        'function (traced): %s\n  as a device-style function (numba CUDA): %s' %
                        (getattr(func, '__name__', kind), trace_err,
                         str(err).splitlines()[0])) from err
    fn.second_order = kind == 'F' and n == 4
    fn.style = 'device'
    return fn


def generate_cfuncs(F, B, S, ndim, V):
    """Same call as reference cfuncs.py:30; returns device-function descriptors."""
    return _lower(F, 'F', ndim, V), _lower(B, 'B', ndim, V), _lower(S, 'S', ndim, V)
