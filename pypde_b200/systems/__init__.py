"""Ready-made PDE systems (the reference's example problems, pypde/tests/*).

Each system exists as one C text (systems_src.h) that NVRTC compiles for the
GPU and gcc compiles for the CPU reference, so both sides evaluate the very
same expressions.
"""
import os

from pypde_b200.cfuncs import CudaSource

_HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (V as a function of ndim, has F, has B, has S, second order)
SYSTEMS = {
    'euler': (lambda nd: 2 + nd, True, False, False, False),
    'reactive_euler': (lambda nd: 3 + nd, True, False, True, False),
    'navier_stokes': (lambda nd: 5, True, False, False, True),
    'advect_nc': (lambda nd: 3, True, True, True, False),
    'gpr': (lambda nd: 17, True, True, True, False),
    'burgers': (lambda nd: 1, True, False, False, False),
}


def header_text():
    with open(os.path.join(_HERE, 'systems_src.h')) as f:
        return f.read()


def cuda_sources(system, ndim, defines=None):
    """Returns (F, B, S) `CudaSource` descriptors (None where the system has no
    such term) and V."""
    vfun, hasF, hasB, hasS, second = SYSTEMS[system]
    text = header_text()
    extra = ''.join('#define %s %s\n' % kv for kv in (defines or {}).items())

    def one(kind):
        # each image defines only its own symbol; the other two get throwaway
        # static names so that nothing is defined twice at link time
        names = {k: ('user_' + k if k == kind else '_unused_%s_%s' % (kind, k))
                 for k in 'FBS'}
        src = ('#define SYS_%s\n#define SYS_NDIM %d\n%s'
               '#define PDE_FN extern "C" __device__\n'
               '#define SYS_F %s\n#define SYS_B %s\n#define SYS_S %s\n' %
               (system.upper(), ndim, extra, names['F'], names['B'],
                names['S'])) + text
        # unused functions must not be exported: make them static inline
        for k in 'FBS':
            if k != kind:
                src = src.replace('PDE_FN void SYS_%s(' % k,
                                  'static __device__ inline void SYS_%s(' % k)
        fn = CudaSource(src, '%s_%dd_%s' % (system, ndim, kind))
        fn.second_order = second
        return fn

    return (one('F') if hasF else None, one('B') if hasB else None,
            one('S') if hasS else None, vfun(ndim))


def euler_wavespeed(ndim):
    """OPT-IN analytic wave speed |v_d| + c of the Euler system above, for
    `pde_solver(..., wavespeed=...)` / `pypde_b200_set_wavespeed` (not the reference's
    finite-difference definition: see include/pypde_b200.h)."""
    return CudaSource(
        'extern "C" __device__ double user_L(const double *Q, const double *dQ, int d) {\n'
        '  const double g = 1.4;\n'
        '  const double r = Q[0], ir = 1. / r;\n'
        '  double vv = 0.;\n'
        '  for (int i = 0; i < %d; i++) { const double v = Q[2 + i] * ir; vv += v * v; }\n'
        '  const double p = (g - 1.) * (Q[1] - 0.5 * r * vv);\n'
        '  return fabs(Q[2 + d] * ir) + sqrt(g * p * ir);\n'
        '}\n' % ndim, 'euler_%dd_L' % ndim)
