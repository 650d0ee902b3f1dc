"""The 2-D Euler flux of BASELINE configs[1] written in Python, in the two styles
`pypde_b200.cfuncs` lowers (bench.py --user-functions numba | traced):

* device style  F(out, Q, d) — scalar code, lowered through numba's CUDA target to LTO-IR;
* reference style  F(Q, d) -> ndarray — exactly what reference pypde.pde_solver accepts
  (pypde/tests/euler/system.py), lowered by symbolic tracing to CUDA source.

Both evaluate the expressions of systems_src.h (SYS_EULER) in the same order, so all three
forms give the same bits.
"""
import numpy as np


def F_euler2d_device(out, Q, d):
    g = 1.4
    r = Q[0]
    ir = 1. / r
    E = Q[1] * ir
    v0 = Q[2] * ir
    v1 = Q[3] * ir
    vv = 0. + v0 * v0
    vv = vv + v1 * v1
    e = E - vv / 2.
    p = (g - 1.) * r * e
    vd = v0 if d == 0 else v1
    out[0] = r * vd
    out[1] = r * E * vd + p * vd
    out[2] = r * v0 * vd
    out[3] = r * v1 * vd
    out[2 + d] += p


def F_euler2d_reference(Q, d):
    g = 1.4
    r = Q[0]
    ir = 1. / r
    E = Q[1] * ir
    v = Q[2:4] * ir
    vv = 0. + v[0] * v[0]
    vv = vv + v[1] * v[1]
    e = E - vv / 2.
    p = (g - 1.) * r * e
    ret = np.zeros(4)
    ret[0] = r * v[d]
    ret[1] = r * E * v[d] + p * v[d]
    ret[2:4] = r * v * v[d]
    ret[2 + d] += p
    return ret


def euler2d(style):
    """The lowered flux (a DeviceFunction) for style 'numba' or 'traced'."""
    from pypde_b200.cfuncs import generate_cfuncs
    f = {'numba': F_euler2d_device, 'traced': F_euler2d_reference}[style]
    return generate_cfuncs(f, None, None, 2, 4)[0]
