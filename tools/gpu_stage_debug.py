"""Stage-by-stage comparison of the CUDA kernels with the reference classes
(oracle/_ref/libpypde_stages.so) on identical inputs.  Diagnostic script; the
assertions live in tests/test_gpu_parity.py."""
import sys
import numpy as np

sys.path.insert(0, '.')
from oracle import reference as R
from pypde_b200.handle import Solver
from pypde_b200.systems import cuda_sources
import ctypes
from pypde_b200.utils import get_cdll


def tables(N):
    lib = get_cdll()
    P = ctypes.POINTER(ctypes.c_double)
    t = {k: np.zeros(s) for k, s in [('nodes', N), ('wghts', N), ('derv', (N, N)), ('endv', (2, N))]}
    lib.pypde_b200_tables(N, t['nodes'].ctypes.data_as(P), t['wghts'].ctypes.data_as(P),
                          t['derv'].ctypes.data_as(P), t['endv'].ctypes.data_as(P), None, None,
                          None, None, None)
    return t


def rel(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


def traces_from_qh(qh, nXw, N, ndim, V, endv):
    ncw = int(np.prod(nXw))
    q = qh.reshape([ncw, N] + [N] * ndim + [V])
    NP = N**ndim
    out = np.zeros((ncw, ndim, 2, NP, V))
    for d in range(ndim):
        for e in range(2):
            tr = np.tensordot(q, endv[e], axes=([2 + d], [0]))  # removes axis 2+d
            out[:, d, e] = tr.reshape(ncw, NP, V)
    return out


def smooth_ic(shape, system, ndim):
    g = 1.4
    grids = np.meshgrid(*[(np.arange(n) + 0.5) / n for n in shape], indexing='ij')
    s = np.ones(shape)
    for x in grids:
        s = s * np.sin(2 * np.pi * x)
    rho = 1 + 0.2 * s
    vel = [1.0, -0.5, 0.25][:ndim]
    p = 1.0
    if system == 'euler':
        Q = np.zeros(tuple(shape) + (2 + ndim, ))
        Q[..., 0] = rho
        Q[..., 1] = p / (g - 1) + rho * sum(v * v for v in vel) / 2
        for i, v in enumerate(vel):
            Q[..., 2 + i] = rho * v
        return Q
    if system == 'advect_nc':
        Q = np.zeros(tuple(shape) + (3, ))
        Q[..., 0] = rho
        Q[..., 1] = 0.5 + 0.1 * s
        Q[..., 2] = 1.0 - 0.3 * s
        return Q
    raise ValueError(system)


def run(system, shape, N, bts, steps=3, cfl=0.9):
    ndim = len(shape)
    F, B, S, V = cuda_sources(system, ndim)
    cF, cB, cS = R.system_callbacks(system, ndim)
    L = [1.0] * ndim
    dX = np.array([L[i] / shape[i] for i in range(ndim)])
    Q0 = smooth_ic(shape, system, ndim)
    tf = 10.0
    sol = Solver(Q0.shape, L, F=F, B=B, S=S, boundaryTypes=bts, cfl=cfl, order=N, stiff=False)
    sol.set_state(Q0)
    sol.begin(tf)
    st = R.Stages()
    tb = tables(N)
    u = Q0.copy()
    t = 0.0
    nXw = [n + 2 for n in shape]
    print('== %s shape=%s N=%d bt=%s' % (system, shape, N, bts))
    for k in range(steps):
        tg, dtg, nan = sol.step()
        ub_ref = st.boundaries(u, bts, N)
        ub = sol.read_stage('ub').reshape(ub_ref.shape)
        w_ref = R.weno_solver(ub_ref, N)
        w = sol.read_stage('w').reshape(w_ref.shape)
        # same-input variant: reference WENO applied to the GPU's ub
        dt_ref = st.step(cF, cB, w_ref, dX, N, cfl, tf, False, t, k)
        qh_ref = st.predictor(cF, cB, cS, w_ref, dX, N, dt_ref)
        tr_ref = traces_from_qh(qh_ref, nXw, N, ndim, V, tb['endv'])
        tr = sol.read_stage('traces').reshape(tr_ref.shape)
        u_ref = st.fv(u, cF, cB, cS, qh_ref, dX, N, dt_ref)
        ug = sol.get_state()
        print(' step %d: ub %.1e  w %.1e  dt %.3e (ref %.3e, rel %.1e)  traces %.1e  u %.2e  nan=%s' %
              (k, rel(ub, ub_ref), rel(w, w_ref), dtg, dt_ref, abs(dtg - dt_ref) / dt_ref,
               rel(tr, tr_ref), rel(ug, u_ref), nan))
        u = u_ref
        t += dt_ref
        # keep the two paths on identical inputs for the next step
        sol.set_state(u)
    sol.close()


if __name__ == '__main__':
    run('euler', (64, ), 2, ['periodic'])
    run('euler', (64, ), 3, ['transitive'])
    run('euler', (24, 20), 2, ['periodic', 'transitive'])
    run('euler', (24, 20), 3, ['periodic', 'periodic'])
    run('advect_nc', (40, ), 3, ['periodic'])
    run('advect_nc', (16, 12), 2, ['periodic', 'periodic'])
    run('euler', (16, 12), 4, ['periodic', 'periodic'])
