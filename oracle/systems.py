"""Test infrastructure, not product code.

numpy statements of the PDE systems of pypde_b200/systems/systems_src.h, with
the same expression order, vectorised over leading axes:
    F(Q[..., V], dQ[..., ndim, V], d) -> [..., V]
    B(Q[..., V], d)                   -> [..., V, V]
    S(Q[..., V])                      -> [..., V]
"""
import numpy as np


def euler(ndim):
    g = 1.4

    def F(Q, dQ, d):
        r = Q[..., 0]
        ir = 1. / r
        E = Q[..., 1] * ir
        v = [Q[..., 2 + i] * ir for i in range(ndim)]
        vv = 0.
        for i in range(ndim):
            vv = vv + v[i] * v[i]
        e = E - vv / 2.
        p = (g - 1.) * r * e
        vd = v[d]
        out = np.empty_like(Q)
        out[..., 0] = r * vd
        out[..., 1] = r * E * vd + p * vd
        for i in range(ndim):
            out[..., 2 + i] = r * v[i] * vd
        out[..., 2 + d] += p
        return out

    return dict(F=F, B=None, S=None, V=2 + ndim, second_order=False)


def reactive_euler(ndim, K0=250., Ea=2.):
    g, Qc, cv = 1.4, 1., 2.5

    def F(Q, dQ, d):
        r = Q[..., 0]
        ir = 1. / r
        E = Q[..., 1] * ir
        v = [Q[..., 2 + i] * ir for i in range(ndim)]
        vv = 0.
        for i in range(ndim):
            vv = vv + v[i] * v[i]
        lam = Q[..., 2 + ndim] * ir
        e = E - vv / 2. - Qc * (lam - 1.)
        p = (g - 1.) * r * e
        vd = v[d]
        out = vd[..., None] * Q
        out[..., 1] += p * vd
        out[..., 2 + d] += p
        return out

    def S(Q):
        r = Q[..., 0]
        ir = 1. / r
        E = Q[..., 1] * ir
        vv = 0.
        for i in range(ndim):
            vi = Q[..., 2 + i] * ir
            vv = vv + vi * vi
        lam = Q[..., 2 + ndim] * ir
        e = E - vv / 2. - Qc * (lam - 1.)
        T = e / cv
        out = np.zeros_like(Q)
        out[..., 2 + ndim] = -r * lam * K0 * np.exp(-Ea / T)
        return out

    return dict(F=F, B=None, S=S, V=3 + ndim, second_order=False)


def navier_stokes(ndim, mu=1e-2):
    g = 1.4

    def F(Q, dQ, d):
        r = Q[..., 0]
        ir = 1. / r
        E = Q[..., 1] * ir
        v = [Q[..., 2 + i] * ir for i in range(3)]
        dr_dx = dQ[..., 0, 0]
        dv_dx = [(dQ[..., 0, 2 + i] - dr_dx * v[i]) * ir for i in range(3)]
        p = r * (g - 1.) * (E - (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / 2.)
        tr = dv_dx[0]
        sd = []
        for j in range(3):
            dv_dj = dv_dx[j] if d == 0 else 0.
            dv_jd = dv_dx[d] if j == 0 else 0.
            I = 1. if d == j else 0.
            sd.append(mu * (dv_dj + dv_jd - 2. / 3. * tr * I))
        vd = v[d]
        rvd = r * vd
        out = np.empty_like(Q)
        out[..., 0] = rvd
        out[..., 1] = rvd * E + p * vd
        for i in range(3):
            out[..., 2 + i] = rvd * v[i]
        out[..., 2 + d] += p
        out[..., 1] -= sd[0] * v[0] + sd[1] * v[1] + sd[2] * v[2]
        for i in range(3):
            out[..., 2 + i] -= sd[i]
        return out

    return dict(F=F, B=None, S=None, V=5, second_order=True)


def advect_nc(ndim):
    def F(Q, dQ, d):
        a = 1. - 0.35 * d
        out = np.zeros_like(Q)
        out[..., 0] = a * Q[..., 0] * (1. + Q[..., 0] / 4.)
        return out

    def B(Q, d):
        a = 0.6 + 0.3 * d
        out = np.zeros(Q.shape + (3, ))
        out[..., 1, 1] = a * (1. + 0.2 * Q[..., 0])
        out[..., 1, 2] = 0.1 * Q[..., 1]
        out[..., 2, 0] = 0.05
        out[..., 2, 2] = a + 0.1 * Q[..., 2]
        return out

    def S(Q):
        out = np.empty_like(Q)
        out[..., 0] = -0.5 * (Q[..., 0] - 1.)
        out[..., 1] = 0.3 * Q[..., 2] - 0.2 * Q[..., 1]
        out[..., 2] = -0.1 * Q[..., 2] * Q[..., 0]
        return out

    return dict(F=F, B=B, S=S, V=3, second_order=False)


SYSTEMS = {
    'euler': euler,
    'reactive_euler': reactive_euler,
    'navier_stokes': navier_stokes,
    'advect_nc': advect_nc,
}
