"""Test infrastructure, not product code.

CPU restatement (numpy, vectorised over cells) of the reference's ADER-WENO
stepping path, used ONLY as the checker in tests/, smoke() and bench.py's
cpu_baseline leg.  Every function cites the reference file:line it follows
(paths relative to the reference's src/).

Pinned: the reference ships no golden vectors for this path, so
this restatement is pinned against the reference itself — built unmodified into
oracle/_ref by oracle/Makefile and run in tests/test_oracle.py — and against
the fixtures under tests/golden/ that tests/golden/make_golden.py generated
from that build.

User functions are numpy callables (oracle/systems.py):
    F(Q[..., V], dQ[..., ndim, V], d), B(Q[..., V], d), S(Q[..., V]).
"""
import numpy as np

LAMS, LAMC, EPS = 1., 1e5, 1e-14   # solvers/weno/weno.h:8-10
DG_IT, DG_TOL = 50, 6e-6           # solvers/dg/dg.h:8-9
SQRT_EPS = np.sqrt(np.finfo(float).eps)  # eigs/NumericalDiff.h:9
RUSANOV, ROE, OSHER = 0, 1, 2      # solvers/fv/fluxes.h:6-8
DBL_MAX = np.finfo(float).max      # types.h:19  (INF is DBL_MAX, not inf)


# ---------------------------------------------------------------------------
# tables: poly/basis.cpp, solvers/weno/weno_matrices.cpp, solvers/dg/dg_matrices.cpp
# ---------------------------------------------------------------------------
class Tables:
    def __init__(self, N):
        self.N = N
        x, w = np.polynomial.legendre.leggauss(N)     # scipy/math/legendre.cpp:136-197
        self.nodes = (x + 1) / 2                      # basis.cpp:7-13
        self.wghts = w / 2                            # basis.cpp:15-20
        # Lagrange basis at the nodes (basis.cpp:22-53)
        self.psi = []
        for i in range(N):
            p = np.poly1d([1.])
            for j in range(N):
                if j != i:
                    p = p * np.poly1d([1., -self.nodes[j]]) / (self.nodes[i] - self.nodes[j])
            self.psi.append(p)
        psi = self.psi
        # basis.cpp:55-76
        self.endv = np.array([[p(0.) for p in psi], [p(1.) for p in psi]])
        self.derv = np.array([[psi[j].deriv(1)(self.nodes[i]) for j in range(N)]
                              for i in range(N)])
        # dg_matrices.cpp:27-55, dg.cpp:36-39
        dg_end = np.array([[psi[i](1.) * psi[j](1.) for j in range(N)] for i in range(N)])
        dg_der = np.zeros((N, N))
        for i in range(N):
            for j in range(N):
                if i == j:
                    dg_der[i, j] = (psi[i](1.)**2 - psi[i](0.)**2) / 2
                else:
                    dg_der[i, j] = self.wghts[i] * psi[j].deriv(1)(self.nodes[i])
        self.dgmat = dg_end - dg_der.T
        # weno_matrices.cpp:8-36 and weno.cpp:9-10
        FN2 = int(np.floor((N - 1) / 2.))
        CN2 = int(np.ceil((N - 1) / 2.))
        first = [-(N - 1), 0, -CN2, -FN2]
        self.wm = []
        for s in range(4):
            m = np.zeros((N, N))
            for i in range(N):
                for j in range(N):
                    P = psi[j].integ()
                    a = first[s] + i
                    m[i, j] = P(a + 1) - P(a)
            self.wm.append(m)
        # weno_matrices.cpp:38-51
        self.sig = np.zeros((N, N))
        for i in range(N):
            for j in range(N):
                for a in range(1, N):
                    P = (psi[i].deriv(a) * psi[j].deriv(a)).integ()
                    self.sig[i, j] += P(1.) - P(0.)
        # stencil windows inside the 2N-1 line, weno.cpp:41-59
        self.stencils = [(0, 0, LAMS), (1, N - 1, LAMS)]
        if N > 2:
            self.stencils.append((2, FN2, LAMC))
            if N % 2 == 0:
                self.stencils.append((3, CN2, LAMC))


_tables = {}


def tables(N):
    if N not in _tables:
        _tables[N] = Tables(N)
    return _tables[N]


# ---------------------------------------------------------------------------
# grid/boundaries.cpp:7-53
# ---------------------------------------------------------------------------
def boundaries(u, boundary_types, N):
    """u: (nX..., V) -> (nX+2N..., V); 0 transmissive = clamp, 1 periodic = wrap."""
    ndim = u.ndim - 1
    ub = u
    for d in range(ndim):
        pad = [(0, 0)] * ub.ndim
        pad[d] = (N, N)
        ub = np.pad(ub, pad, mode='edge' if boundary_types[d] == 0 else 'wrap')
    return ub


# ---------------------------------------------------------------------------
# solvers/weno/weno.cpp:26-121
# ---------------------------------------------------------------------------
def weno(ub, N, ndim=None):
    """ub: (m_0..m_{n-1}, V) -> (m_0-2(N-1).., N,..,N, V); sweeps d = 0 first
    (weno.cpp:70); the new nodal axis of sweep d goes after the existing ones."""
    T = tables(N)
    if ndim is None:
        ndim = ub.ndim - 1
    rec = ub
    for d in range(ndim):
        # windows of 2N-1 cells along axis d (weno.cpp:107-112)
        win = np.lib.stride_tricks.sliding_window_view(rec, 2 * N - 1, axis=d)
        num = 0.
        den = 0.
        for (s, off, lam) in T.stencils:
            data = win[..., off:off + N]
            # c = M^-1 data (weno.cpp:28)
            c = np.linalg.solve(T.wm[s], data[..., None])[..., 0]
            # p = c^T SIG c + EPS ; o = LAM / p^8 (weno.cpp:31-35)
            p = np.einsum('...a,ab,...b->...', c, T.sig, c) + EPS
            p2 = p * p
            p4 = p2 * p2
            p8 = p4 * p4
            o = lam / p8
            num = num + o[..., None] * c
            den = den + o
        out = num / den[..., None]          # weno.cpp:57 ; shape (..., V, N)
        rec = np.moveaxis(out, -1, -2)      # (cells..., a_0..a_d, V)
    return rec


# ---------------------------------------------------------------------------
# poly/evaluations.cpp:23-70  (with the intended indexing; identical to the
# reference for ndim <= 2 and to reference + the zero_index fix for ndim = 3)
# ---------------------------------------------------------------------------
def derivs(q, d, ndim, dX, T):
    """q: (..., [N]*ndim, V); derivative along nodal axis d, divided by dX[d]."""
    ax = q.ndim - 1 - ndim + d
    r = np.tensordot(q, T.derv, axes=([ax], [1]))   # new axis last: index a
    r = np.moveaxis(r, -1, ax)
    return r / dX[d]


def endpts(q, d, e, ndim, T):
    """q: (..., [N]*ndim, V) -> (..., [N]*(ndim-1), V): trace at end e of axis d."""
    ax = q.ndim - 1 - ndim + d
    return np.tensordot(q, T.endv[e], axes=([ax], [0]))


def weight_products(ndim, T):
    wp = np.ones([T.N] * ndim)
    for d in range(ndim):
        shape = [1] * ndim
        shape[d] = T.N
        wp = wp * T.wghts.reshape(shape)
    return wp


# ---------------------------------------------------------------------------
# eigs/system.cpp:6-61, eigs/NumericalDiff.h:6-44
# ---------------------------------------------------------------------------
def fd_jacobian(F, q, dq, d, second_order_var=False):
    """Forward differences, h = max(eps|x|, eps); q: (n, V), dq: (n, ndim, V).
    second_order_var: differentiate w.r.t. dq[:, d, :] instead of q."""
    n, V = q.shape
    f0 = F(q, dq, d)
    J = np.empty((n, V, V))
    for i in range(V):
        q2, dq2 = q.copy(), dq.copy()
        x = dq2[:, d, i] if second_order_var else q2[:, i]
        h = np.maximum(SQRT_EPS * np.abs(x), SQRT_EPS)
        x += h
        f1 = F(q2, dq2, d)
        J[:, :, i] = (f1 - f0) / h[:, None]
    return J


def system_matrix(F, B, q, dq, d):
    n, V = q.shape
    M = fd_jacobian(F, q, dq, d) if F is not None else np.zeros((n, V, V))
    if B is not None:
        M = M + B(q, d)
    return M


def max_abs_eig(M):
    # system.cpp:28-43: Eigen EigenSolver (V<6) / Spectra with ncv=V: spectral radius
    return np.abs(np.linalg.eigvals(M)).max(axis=-1)


def max_abs_eigs(F, B, q, dq, d):
    return max_abs_eig(system_matrix(F, B, q, dq, d))


def max_abs_eigs_second_order(F, q, dq, d, N, dX):
    return 2 * (N + 1) / dX[d] * max_abs_eig(fd_jacobian(F, q, dq, d, True))


# ---------------------------------------------------------------------------
# solvers/stepper.cpp:35-76
# ---------------------------------------------------------------------------
def cfl_max(w, F, B, dX, N, second_order):
    """max over every cell of w (ghost layer included) of sum_d lambda_d/dx_d."""
    T = tables(N)
    ndim = len(dX)
    V = w.shape[-1]
    wp = weight_products(ndim, T)
    nodal = tuple(range(w.ndim - 1 - ndim, w.ndim - 1))
    wc = w.reshape((-1, ) + w.shape[-1 - ndim:])
    nod = tuple(range(1, 1 + ndim))
    q = np.tensordot(wc, wp, axes=(nod, tuple(range(ndim))))
    dq = np.stack([np.tensordot(derivs(wc, d, ndim, dX, T), wp, axes=(nod, tuple(range(ndim))))
                   for d in range(ndim)], axis=1)
    tmp = np.zeros(q.shape[0])
    for d in range(ndim):
        lam = max_abs_eigs(F, B, q, dq, d)
        if second_order:
            lam = lam + max_abs_eigs_second_order(F, q, dq, d, N, dX)
        tmp = tmp + lam / dX[d]
    # std::max(MAX, tmp) drops NaNs
    tmp = tmp[~np.isnan(tmp)]
    return max(0., tmp.max()) if tmp.size else 0.


def time_step(w, F, B, dX, N, cfl, tf, second_order, t, count):
    MAX = cfl_max(w, F, B, dX, N, second_order)
    with np.errstate(divide='ignore'):
        dt = np.float64(cfl) / np.float64(MAX)
    if count <= 5:
        dt *= 0.2
    if t + dt > tf:
        return tf - t
    return float(dt)


# ---------------------------------------------------------------------------
# solvers/dg/dg.cpp:50-126 (rhs), 156-171, 187-224 (Picard predictor)
# ---------------------------------------------------------------------------
def dg_rhs(q, Ww, dt, F, B, S, dX, N, exact_b=False):
    """q: (n, N_t, [N]*ndim, V)."""
    T = tables(N)
    ndim = len(dX)
    wp = weight_products(ndim, T)
    ret = np.empty_like(q)
    for t in range(N):
        qt = q[:, t]
        dq = np.stack([derivs(qt, d, ndim, dX, T) for d in range(ndim)], axis=-2)
        r = S(qt) if S is not None else np.zeros_like(qt)
        for d in range(ndim):
            if B is not None:
                b = B(qt, d)
                if exact_b:
                    r = r - np.einsum('...ij,...j->...i', b, dq[..., d, :])
                else:
                    # dg.cpp:109-110: `b * dq.row(k)` (V x V times 1 x V) evaluates,
                    # with Eigen's checks compiled out, as b(0,0) * dq_row
                    r = r - b[..., 0, 0][..., None] * dq[..., d, :]
            if F is not None:
                f = F(qt, dq, d)
                r = r - derivs(f, d, ndim, dX, T)
        c = T.wghts[t]
        cw = c * np.ones([N] * ndim)
        for d in range(ndim):       # dg.cpp:104,117: c = w_t, then *= w_a per dim
            shape = [1] * ndim
            shape[d] = N
            cw = cw * T.wghts.reshape(shape)
        ret[:, t] = r * cw[..., None]
    return ret * dt + Ww


def dg_initial_condition(w, N, ndim):
    T = tables(N)
    Ww = np.empty((w.shape[0], N) + w.shape[1:])
    for t in range(N):
        c = T.endv[0, t] * np.ones([N] * ndim)
        for d in range(ndim):       # dg.cpp:162-164
            shape = [1] * ndim
            shape[d] = N
            c = c * T.wghts.reshape(shape)
        Ww[:, t] = c[..., None] * w
    return Ww


def predictor(w, dt, F, B, S, dX, N, exact_b=False):
    """w: (cells..., [N]*ndim, V) -> qh: (ncell, N_t, [N]*ndim, V). Non-stiff."""
    T = tables(N)
    ndim = len(dX)
    wc = w.reshape((-1, ) + w.shape[-1 - ndim:])
    n = wc.shape[0]
    wp = weight_products(ndim, T)
    dginv = np.linalg.inv(T.dgmat)
    Ww = dg_initial_condition(wc, N, ndim)
    q0 = np.repeat(wc[:, None], N, axis=1)          # dg.cpp:12-20
    q1 = q0.copy()
    active = np.arange(n)
    for _ in range(DG_IT):
        if active.size == 0:
            break
        qa = q0[active]
        rhs = dg_rhs(qa, Ww[active], dt, F, B, S, dX, N, exact_b)
        # DG_U = kron(DG_MAT, diag W, ..) (dg.cpp:42-47): solve along the time index
        qn = np.einsum('tk,nk...->nt...', dginv, rhs) / wp[..., None]
        q1[active] = qn
        diff = np.abs(qn - qa) > DG_TOL * (1. + np.abs(qa))
        changed = diff.reshape(diff.shape[0], -1).any(axis=1)
        q0[active[changed]] = qn[changed]
        active = active[changed]
    return q1


# ---------------------------------------------------------------------------
# solvers/fv/fluxes.cpp:20-117
# ---------------------------------------------------------------------------
def _abs_matrix_apply(M, Dq):
    """Re(R |Lambda| R^-1 Dq), fluxes.cpp:36-41,64-69."""
    lam, R = np.linalg.eig(M)
    b = np.linalg.solve(R, Dq.astype(complex)[..., None])[..., 0] * np.abs(lam)
    return np.einsum('...ij,...j->...i', R, b).real


def interface_flux(F, B, qL, qR, dqL, dqR, d, N, dX, flux, second_order):
    T = tables(N)
    if flux == RUSANOV:
        m = np.maximum(max_abs_eigs(F, B, qL, dqL, d), max_abs_eigs(F, B, qR, dqR, d))
        ret = m[:, None] * (qL - qR)
    else:
        Dq = qL - qR
        Ddq = dqL - dqR
        if flux == ROE:
            M = 0.
            for i in range(N):
                q = qR + T.nodes[i] * Dq
                dq = dqR + T.nodes[i] * Ddq
                M = M + T.wghts[i] * system_matrix(F, B, q, dq, d)
            ret = _abs_matrix_apply(M, Dq)
        else:
            ret = 0.
            for i in range(N):
                q = qR + T.nodes[i] * Dq
                dq = dqR + T.nodes[i] * Ddq
                ret = ret + T.wghts[i] * _abs_matrix_apply(system_matrix(F, B, q, dq, d), Dq)
    ret = ret + F(qL, dqL, d) + F(qR, dqR, d)
    if second_order:
        m = np.maximum(max_abs_eigs_second_order(F, qL, dqL, d, N, dX),
                       max_abs_eigs_second_order(F, qR, dqR, d, N, dX))
        ret = ret + m[:, None] * (qL - qR)
    return ret


def b_path_integral(B, qL, qR, d, N):
    T = tables(N)
    Dq = qR - qL
    M = 0.
    for i in range(N):
        M = M + T.wghts[i] * B(qL + T.nodes[i] * Dq, d)
    return np.einsum('...ij,...j->...i', M, Dq)


# ---------------------------------------------------------------------------
# solvers/fv/fv.cpp:33-207
# ---------------------------------------------------------------------------
def fv_apply(u, qh, dt, F, B, S, dX, N, flux=RUSANOV, second_order=False, exact_b=False):
    """u: (nX..., V) (returned updated); qh: (nX+2..., N_t, [N]*ndim, V)."""
    T = tables(N)
    ndim = len(dX)
    V = u.shape[-1]
    nX = u.shape[:-1]
    u = u.copy()
    wp = weight_products(ndim, T)
    inner = tuple(slice(1, -1) for _ in range(ndim))

    if B is not None or S is not None:       # centers, fv.cpp:33-85
        qi = qh[inner]
        for t in range(N):
            qt = qi[(slice(None), ) * ndim + (t, )]
            s = S(qt) if S is not None else np.zeros_like(qt)
            if B is not None:
                for d in range(ndim):
                    dq = derivs(qt, d, ndim, dX, T)
                    b = B(qt, d)
                    if exact_b:
                        s = s - np.einsum('...ij,...j->...i', b, dq)
                    else:
                        # fv.cpp:68-70: column destination, inner size 1:
                        # s(i) -= dq_row(0) * b(i,0)
                        s = s - dq[..., 0][..., None] * b[..., :, 0]
            c = dt * T.wghts[t] * wp
            nod = tuple(range(ndim, 2 * ndim))
            u += (c[..., None] * s).sum(axis=nod)

    if F is not None or B is not None:       # interfaces, fv.cpp:124-198
        for d in range(ndim):
            # faces between w-cells i_d and i_d+1, i_d = 0..nX_d; transverse interior
            sl_L = [slice(1, -1)] * ndim
            sl_R = [slice(1, -1)] * ndim
            sl_L[d] = slice(0, -1)
            sl_R[d] = slice(1, None)
            qL_c = qh[tuple(sl_L)]
            qR_c = qh[tuple(sl_R)]
            fshape = qL_c.shape[:ndim]
            wpt = weight_products(ndim - 1, T) if ndim > 1 else np.ones(())
            FL = np.zeros(fshape + (V, ))
            FR = np.zeros(fshape + (V, ))
            for t in range(N):
                idx = (slice(None), ) * ndim + (t, )
                q0 = endpts(qL_c[idx], d, 1, ndim, T)     # fv.cpp:106-107
                q1 = endpts(qR_c[idx], d, 0, ndim, T)
                tshape = q0.shape[:-1]
                if second_order:
                    dq0 = np.stack([endpts(derivs(qL_c[idx], k, ndim, dX, T), d, 1, ndim, T)
                                    for k in range(ndim)], axis=-2)
                    dq1 = np.stack([endpts(derivs(qR_c[idx], k, ndim, dX, T), d, 0, ndim, T)
                                    for k in range(ndim)], axis=-2)
                else:
                    dq0 = np.zeros(tshape + (ndim, V))
                    dq1 = np.zeros(tshape + (ndim, V))
                qLf, qRf = q0.reshape(-1, V), q1.reshape(-1, V)
                dqLf, dqRf = dq0.reshape(-1, ndim, V), dq1.reshape(-1, ndim, V)
                f = np.zeros_like(qLf)
                b = np.zeros_like(qLf)
                if F is not None:
                    f = interface_flux(F, B, qLf, qRf, dqLf, dqRf, d, N, dX, flux, second_order)
                if B is not None:
                    b = b_path_integral(B, qLf, qRf, d, N)
                c = dt * T.wghts[t] / (2. * dX[d])          # fv.cpp:159
                cw = (c * wpt)[..., None]                   # fv.cpp:181-183
                tn = tuple(range(ndim, 2 * ndim - 1))
                FL += (cw * (b + f).reshape(tshape + (V, ))).sum(axis=tn)
                FR += (cw * (b - f).reshape(tshape + (V, ))).sum(axis=tn)
            lo = [slice(None)] * ndim
            hi = [slice(None)] * ndim
            lo[d] = slice(1, None)     # faces whose LEFT cell is interior
            hi[d] = slice(0, -1)       # faces whose RIGHT cell is interior
            u -= FL[tuple(lo)]         # fv.cpp:185-186
            u -= FR[tuple(hi)]         # fv.cpp:188-189
    return u


# ---------------------------------------------------------------------------
# solvers/iterator.cpp:38-151
# ---------------------------------------------------------------------------
def step(u, t, count, tf, dX, boundary_types, F, B, S, N, cfl, flux=RUSANOV,
         second_order=False, exact_b=False, stages=None, stiff=False):
    """One time step; returns (u_new, dt)."""
    ndim = len(dX)
    ub = boundaries(u, boundary_types, N)
    w = weno(ub, N, ndim)
    dt = time_step(w, F, B, dX, N, cfl, tf, second_order, t, count)
    if stiff:
        qh = predictor_stiff(w, dt, F, B, S, dX, N, exact_b)
    else:
        qh = predictor(w, dt, F, B, S, dX, N, exact_b)
    qh = qh.reshape(w.shape[:ndim] + qh.shape[1:])
    un = fv_apply(u, qh, dt, F, B, S, dX, N, flux, second_order, exact_b)
    if stages is not None:
        stages.update(ub=ub, w=w, dt=dt, qh=qh)
    return un, dt


def pde_solver(Q0, tf, L, F=None, B=None, S=None, boundary_types=None, cfl=0.9, order=2,
               ndt=100, flux=RUSANOV, second_order=False, max_steps=None, exact_b=False,
               stiff=False):
    """Restates iterator.cpp:99-150.  Returns (ret, nsteps)."""
    u = np.array(Q0, dtype=float)
    ndim = u.ndim - 1
    nX = u.shape[:-1]
    dX = np.array([L[i] / nX[i] for i in range(ndim)])
    if boundary_types is None:
        boundary_types = [0] * ndim
    ret = np.zeros((ndt, ) + u.shape)
    t, count, push = 0., 0, 0
    while t < tf:
        u, dt = step(u, t, count, tf, dX, boundary_types, F, B, S, order, cfl, flux,
                     second_order, exact_b, stiff=stiff)
        t += dt
        count += 1
        if t >= (push + 1) / ndt * tf and push < ndt:
            ret[push] = u
            push += 1
        if np.isnan(u).any():
            break
        if max_steps is not None and count >= max_steps:
            break
    ret[ndt - 1] = u
    return ret, count


# ---------------------------------------------------------------------------
# Stiff predictor: scipy/algos/newton_krylov.cpp:9-183, scipy/algos/lgmres.cpp:6-104,
# solvers/dg/dg.cpp:128-185.  Per cell, plain loops (small cases only).
# ---------------------------------------------------------------------------
MEPS = 2.2204460492503131e-16   # types.h:18


def lgmres(matvec, b, outer_v, tol, maxiter=1, inner_m=30, outer_k=10):
    """lgmres.cpp:6-104 with psolve = identity, x0 = 0.  outer_v is updated in
    place (the recycled augmentation vectors persist across Newton iterations)."""
    n = b.size
    x = np.zeros(n)
    b_norm = np.linalg.norm(b)
    if b_norm == 0.:
        b_norm = 1.
    for _ in range(maxiter):
        r_outer = matvec(x) - b
        r_norm = np.linalg.norm(r_outer)
        if r_norm <= tol * b_norm or r_norm <= tol:
            break
        vs0 = -r_outer
        inner_res_0 = np.linalg.norm(vs0)
        vs0 = vs0 / inner_res_0
        vs = [vs0]
        ws = []
        ind = 1 + inner_m + len(outer_v)
        H = np.zeros((ind + 1, ind))          # Hessenberg-like matrix, column j-1 = hcur
        Q = None
        R = None
        j = 1
        while j < ind:
            if j < len(outer_v) + 1:
                z = outer_v[j - 1]
            elif j == len(outer_v) + 1:
                z = vs0
            else:
                z = vs[-1]
            v_new = matvec(z)
            v_new_norm = np.linalg.norm(v_new)
            hcur = np.zeros(j + 1)
            for i, v in enumerate(vs):
                alpha = np.dot(v, v_new)
                hcur[i] = alpha
                v_new = v_new - alpha * v
            hcur[j] = np.linalg.norm(v_new)
            v_new = v_new / hcur[j]
            vs.append(v_new)
            ws.append(z)
            H[:j + 1, j - 1] = hcur
            # lgmres.cpp:70-77 re-factorises [Q R | hcur] by Householder QR; the
            # least-squares quantities below do not depend on how QR is computed
            Q, R = np.linalg.qr(H[:j + 1, :j], mode='complete')
            inner_res = abs(Q[0, j]) * inner_res_0
            if inner_res <= tol * inner_res_0 or hcur[j] <= MEPS * v_new_norm:
                break
            j += 1
        if j == ind:
            j -= 1
        # lgmres.cpp:86-88: y = R(0:j,0:j)^-1 Q(0,0:j)^T * inner_res_0
        y = np.linalg.solve(R[:j, :j], Q[0, :j]) * inner_res_0
        dx = y[0] * ws[0]
        for i in range(1, j):
            dx = dx + y[i] * ws[i]
        nx = np.linalg.norm(dx)
        if nx > 0.:
            outer_v.append(dx / nx)
        while len(outer_v) > outer_k:
            outer_v.pop(0)
        x = x + dx
    return x


def _armijo(phi, phi0):
    """newton_krylov.cpp:91-142 (scalar_search_armijo with c1 = 1e-4, amin = 1e-2)."""
    c1, amin = 1e-4, 1e-2
    phi_a0 = phi(1.)
    if phi_a0 <= phi0 - c1 * phi0:
        return 1.
    alpha1 = phi0 / (2. * phi_a0)
    phi_a1 = phi(alpha1)
    if phi_a1 <= phi0 - c1 * alpha1 * phi0:
        return alpha1
    while alpha1 > amin:
        factor = alpha1 * alpha1 * (alpha1 - 1)
        a = phi_a1 - phi0 + phi0 * alpha1 - alpha1 * alpha1 * phi_a0
        a /= factor
        b = -(phi_a1 - phi0 + phi0 * alpha1) + alpha1 * alpha1 * alpha1 * phi_a0
        b /= factor
        alpha2 = (-b + np.sqrt(abs(b * b + 3 * a * phi0))) / (3. * a)
        phi_a2 = phi(alpha2)
        if phi_a2 <= phi0 - c1 * alpha2 * phi0:
            return alpha2
        if (alpha1 - alpha2) > alpha1 / 2.0 or (1 - alpha2 / alpha1) < 0.96:
            alpha2 = alpha1 / 2.0
        alpha1 = alpha2
        phi_a0 = phi_a1
        phi_a1 = phi_a2
    return 1.


def nonlin_solve(func, x, f_tol, counters=None):
    """newton_krylov.cpp:144-183 with the defaults f_rtol = x_tol = x_rtol = DBL_MAX.
    The outer `dx` is never updated (the inner one shadows it, :153,:167), so the
    termination test (:61-76) reduces to  ||F||inf <= f_tol AND ||x||inf >= 1,
    or ||F||inf == 0."""
    x = x.copy()
    gamma, eta_max, eta_treshold, eta = 0.9, 0.9999, 0.1, 1e-3
    Fx = func(x)
    Fx_norm = np.abs(Fx).max()            # maxnorm
    rdiff = MEPS**0.5
    outer_v = []
    f0_norm = [0.]

    def check(f, xx):
        f_norm = np.abs(f).max()
        x_norm = np.abs(xx).max()
        dx_norm = DBL_MAX
        if f0_norm[0] == 0.:
            f0_norm[0] = f_norm
        if f_norm == 0.:
            return True
        return (f_norm <= f_tol and f_norm / DBL_MAX <= f0_norm[0]) and \
               (dx_norm <= DBL_MAX and dx_norm / DBL_MAX <= x_norm)

    x0, f0 = x.copy(), Fx.copy()
    maxiter = 3 * (x.size + 1)
    nit = 0
    for _ in range(maxiter):
        if check(Fx, x):
            break
        nit += 1
        omega = rdiff * max(1., np.abs(x0).max()) / max(1., np.abs(f0).max())

        def matvec(v):
            nv = np.linalg.norm(v)
            if nv == 0.:
                return 0. * v
            sc = omega / nv
            return (func(x0 + sc * v) - f0) / sc

        tol = min(eta, eta * Fx_norm)
        dx = -lgmres(matvec, Fx, outer_v, tol)
        # _nonlin_line_search, newton_krylov.cpp:125-142
        cache = {'s': 0., 'phi': float(np.dot(Fx, Fx)), 'F': Fx}

        def phi(s):
            if s == cache['s']:
                return cache['phi']
            v = func(x + s * dx)
            cache.update(s=s, phi=float(np.dot(v, v)), F=v)
            return cache['phi']

        s = _armijo(phi, cache['phi'])
        x = x + s * dx
        Fx = cache['F'] if s == cache['s'] else func(x)
        Fx_norm_new = np.linalg.norm(Fx)       # note: 2-norm here (:170), max-norm at :155
        x0, f0 = x.copy(), Fx.copy()
        eta_A = gamma * Fx_norm_new * Fx_norm_new / (Fx_norm * Fx_norm)
        if gamma * eta * eta < eta_treshold:
            eta = min(eta_max, eta_A)
        else:
            eta = min(eta_max, max(eta_A, gamma * eta * eta))
        Fx_norm = Fx_norm_new
    if counters is not None:
        counters.append(nit)
    return x


def predictor_stiff(w, dt, F, B, S, dX, N, exact_b=False, counters=None):
    """dg.cpp:128-185,204-224 with STIFF = true: per cell Newton-Krylov on
    obj(q) = rhs(q) - (DG_MAT (x) W) q."""
    T = tables(N)
    ndim = len(dX)
    wc = w.reshape((-1, ) + w.shape[-1 - ndim:])
    ncell = wc.shape[0]
    wp = weight_products(ndim, T)
    out = np.empty((ncell, N) + wc.shape[1:])
    for cidx in range(ncell):
        wi = wc[cidx:cidx + 1]
        Ww = dg_initial_condition(wi, N, ndim)
        q0 = np.repeat(wi[:, None], N, axis=1)
        shape = q0.shape

        def obj(qv):
            q = qv.reshape(shape)
            tmp = dg_rhs(q, Ww, dt, F, B, S, dX, N, exact_b)
            # dg.cpp:135-151: tmp(t) -= DG_MAT(t,k) * prod(w) * q(k)
            tmp = tmp - np.einsum('tk,nk...->nt...', T.dgmat, q) * wp[..., None]
            return tmp.ravel()

        res = nonlin_solve(obj, q0.ravel(), DG_TOL, counters)
        out[cidx] = res.reshape(shape)[0]
    return out
