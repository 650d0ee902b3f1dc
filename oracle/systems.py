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


def gpr(ndim):
    """reference tests/gpr/system.py with params.py (stiffened gas, pINF = 0)."""
    G_, CV, CS2, CA2, MU, PR, RHO0 = 1.4, 2.5, 25., 25., 2e-2, 0.75, 1.
    P0 = 1. / G_
    KAPPA = MU * G_ * CV / PR
    T0 = P0 / (RHO0 * (G_ - 1.) * CV)
    TAU1 = 6. * MU / (RHO0 * CS2)
    TAU2 = KAPPA * RHO0 / (T0 * CA2)

    def unpack(Q):
        r = Q[..., 0]
        ir = 1. / r
        E = Q[..., 1] * ir
        v = Q[..., 2:5] * ir[..., None]
        A = Q[..., 5:14].reshape(Q.shape[:-1] + (3, 3))
        J = Q[..., 14:17] * ir[..., None]
        Gm = np.einsum('...ki,...kj->...ij', A, A)
        tr3 = (Gm[..., 0, 0] + Gm[..., 1, 1] + Gm[..., 2, 2]) / 3.
        dev = Gm - tr3[..., None, None] * np.eye(3)
        devG2 = (dev * dev).sum(axis=(-1, -2))
        psi = CS2 * np.einsum('...ik,...kj->...ij', A, dev)
        E3 = (v * v).sum(axis=-1) / 2.
        E1 = E - E3 - CS2 / 4. * devG2 - CA2 / 2. * (J * J).sum(axis=-1)
        p = E1 * r * (G_ - 1.)
        T = p / (r * (G_ - 1.) * CV)
        return r, E, v, A, J, psi, p, T

    def F(Q, dQ, d):
        r, E, v, A, J, psi, p, T = unpack(Q)
        sig = -r[..., None, None] * np.einsum('...ki,...kj->...ij', A, psi)
        vd = v[..., d]
        rvd = r * vd
        out = np.zeros_like(Q)
        out[..., 0] = rvd
        out[..., 1] = rvd * E + p * vd
        out[..., 2:5] = rvd[..., None] * v
        out[..., 2 + d] += p
        out[..., 1] -= (sig[..., d, :] * v).sum(axis=-1)
        out[..., 2:5] -= sig[..., d, :]
        Av = np.einsum('...ik,...k->...i', A, v)
        for i in range(3):
            out[..., 5 + 3 * i + d] = Av[..., i]
        out[..., 1] += CA2 * J[..., d] * T
        out[..., 14:17] = rvd[..., None] * J
        out[..., 14 + d] += T
        return out

    def B(Q, d):
        v = Q[..., 2:5] / Q[..., 0:1]
        out = np.zeros(Q.shape + (17, ))
        for i in range(5, 14):
            out[..., i, i] = v[..., d]
        for k in range(3):
            out[..., 5 + d, 5 + d + k] -= v[..., k]
            out[..., 8 + d, 8 + d + k] -= v[..., k]
            out[..., 11 + d, 11 + d + k] -= v[..., k]
        return out

    def S(Q):
        r, E, v, A, J, psi, p, T = unpack(Q)
        det = np.linalg.det(A)
        th1 = 3. * det**(5. / 3.) / (CS2 * TAU1)
        th2 = 1. / (CA2 * TAU2 * (r / RHO0) * (T0 / T))
        out = np.zeros_like(Q)
        out[..., 5:14] = -psi.reshape(Q.shape[:-1] + (9, )) * th1[..., None]
        out[..., 14:17] = -r[..., None] * (CA2 * J) * th2[..., None]
        return out

    return dict(F=F, B=B, S=S, V=17, second_order=False)


def burgers(ndim):
    def F(Q, dQ, d):
        a = 1. - 0.4 * d
        return a * Q * Q / 2.

    return dict(F=F, B=None, S=None, V=1, second_order=False)


SYSTEMS = {
    'gpr': gpr,
    'burgers': burgers,
    'euler': euler,
    'reactive_euler': reactive_euler,
    'navier_stokes': navier_stokes,
    'advect_nc': advect_nc,
}
