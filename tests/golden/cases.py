"""Initial data and parameters of the parity cases (shared by the golden
generator, the CPU oracle tests and the GPU parity tests).  Deterministic,
no RNG except where a seed is given."""
import numpy as np

G = 1.4


def euler_state(rho, p, vel):
    """Q = [rho, rho E, rho v...] (reference tests/euler: E = p/((g-1) rho) + v^2/2)."""
    rho = np.asarray(rho, dtype=float)
    Q = np.zeros(rho.shape + (2 + len(vel), ))
    Q[..., 0] = rho
    Q[..., 1] = p / (G - 1) + rho * sum(np.asarray(v) * v for v in vel) / 2
    for i, v in enumerate(vel):
        Q[..., 2 + i] = rho * v
    return Q


def centres(shape, rows=None):
    """Cell-centre coordinates in [0, 1]^ndim; rows = (r0, r1) keeps only that range of
    axis 0 (one rank's slab of the grid `shape`)."""
    ax = [(np.arange(n) + 0.5) / n for n in shape]
    if rows is not None:
        ax[0] = ax[0][rows[0]:rows[1]]
    return np.meshgrid(*ax, indexing='ij')


def smooth_product(shape, rows=None):
    xs = centres(shape, rows)
    s = np.ones(xs[0].shape)
    for x in xs:
        s = s * np.sin(2 * np.pi * x)
    return s


def sod(n=200):
    """BASELINE config 1: Sod shock tube."""
    x = (np.arange(n) + 0.5) / n
    rho = np.where(x < 0.5, 1.0, 0.125)
    p = np.where(x < 0.5, 1.0, 0.1)
    return euler_state(rho, p, [np.zeros(n)])


def euler_smooth(shape, rows=None):
    """The well-conditioned parity IC of SURVEY 7.3-H1."""
    nd = len(shape)
    rho = 1 + 0.2 * smooth_product(shape, rows)
    vel = [1.0, -0.5, 0.25][:nd]
    return euler_state(rho, 1.0, [v * np.ones(rho.shape) for v in vel])


def euler_explosion(shape, rows=None):
    """BASELINE config 2 IC: cylindrical / spherical explosion."""
    r2 = sum((x - 0.5)**2 for x in centres(shape, rows))
    inside = r2 < 0.2**2
    rho = np.where(inside, 1.0, 0.125)
    p = np.where(inside, 1.0, 0.1)
    return euler_state(rho, p, [np.zeros(rho.shape)] * len(shape))


def advect_nc_smooth(shape):
    s = smooth_product(shape)
    Q = np.zeros(tuple(shape) + (3, ))
    Q[..., 0] = 1 + 0.2 * s
    Q[..., 1] = 0.5 + 0.1 * s
    Q[..., 2] = 1.0 - 0.3 * s
    return Q


def taylor_green(shape, rows=None):
    """BASELINE config 5 IC on [0, 2 pi]^3."""
    x, y, z = [2 * np.pi * c for c in centres(shape, rows)]
    rho = np.ones(x.shape)
    v = [np.sin(x) * np.cos(y) * np.cos(z), -np.cos(x) * np.sin(y) * np.cos(z),
         np.zeros(x.shape)]
    p = 100 / G + (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2) / 16
    return euler_state(rho, p, v)


def ns_smooth(shape):
    """Navier-Stokes state (V = 5, three velocity components in any ndim)."""
    rho = 1 + 0.2 * smooth_product(shape)
    vel = [1.0 * np.ones(shape), -0.5 * np.ones(shape), 0.25 + 0.1 * smooth_product(shape)]
    Q = np.zeros(tuple(shape) + (5, ))
    Q[..., 0] = rho
    Q[..., 1] = 1.0 / (G - 1) + rho * sum(v * v for v in vel) / 2
    for i, v in enumerate(vel):
        Q[..., 2 + i] = rho * v
    return Q


def reactive_state(rho, p, vel, lam):
    """Q = [rho, rho E, rho v.., rho lambda], E = p/((g-1) rho) + v^2/2 + Qc (lambda - 1)."""
    rho = np.asarray(rho, dtype=float)
    nd = len(vel)
    Q = np.zeros(rho.shape + (3 + nd, ))
    Q[..., 0] = rho
    Q[..., 1] = p / (G - 1) + rho * sum(np.asarray(v) * v for v in vel) / 2 + rho * (lam - 1.)
    for i, v in enumerate(vel):
        Q[..., 2 + i] = rho * v
    Q[..., 2 + nd] = rho * lam
    return Q


def reactive_disc(shape, smooth=False, rows=None):
    """BASELINE config 3 IC: burnt disc (rho, p, lambda) = (2.8, 2.0, 0) in
    unburnt gas (2.0, 0.8, 1); every state has max|Q| well above 1, as the reference's
    Newton termination rule needs.  smooth=True blends the two states with a tanh."""
    r = np.sqrt(sum((x - 0.5)**2 for x in centres(shape, rows)))
    s = 0.5 * (1 - np.tanh((r - 0.25) / 0.08)) if smooth else (r < 0.25).astype(float)
    rho = 2.0 + 0.8 * s
    p = 0.8 + 1.2 * s
    lam = 1.0 - s
    return reactive_state(rho, p, [np.zeros(r.shape)] * len(shape), lam)


def gpr_state(rho, p, vel):
    """reference tests/gpr/misc/utils.py Cvec with A = rho^(1/3) I, J = 0
    (E = p/((g-1) rho) + v.v/2: the distortion energy of A = c I vanishes)."""
    rho = np.asarray(rho, dtype=float)
    Q = np.zeros(rho.shape + (17, ))
    Q[..., 0] = rho
    Q[..., 1] = p / (G - 1) + rho * sum(np.asarray(v) * v for v in vel) / 2
    for i, v in enumerate(vel):
        Q[..., 2 + i] = rho * v
    for i in range(3):
        Q[..., 5 + 4 * i] = rho**(1. / 3.)
    return Q


def gpr_disc(shape, smooth=True, rows=None):
    """BASELINE config 4 IC: (rho, p) = (4, 4/g) in (2, 2/g)."""
    r = np.sqrt(sum((x - 0.5)**2 for x in centres(shape, rows)))
    s = 0.5 * (1 - np.tanh((r - 0.25) / 0.1)) if smooth else (r < 0.25).astype(float)
    rho = 2.0 + 2.0 * s
    return gpr_state(rho, rho / G, [0.1 * np.ones(r.shape), np.zeros(r.shape),
                                    np.zeros(r.shape)])


def weno_kat_input():
    return np.array([1, 2, 4, 7, 11, 16, 22.]).reshape(7, 1)


def weno_random(shape, seed=7):
    return np.random.default_rng(seed).standard_normal(shape) + 2.0


# name -> dict(system, Q0, tf, L, order, bts, flux, stiff, [defines])
def solver_cases():
    c = {}
    c['sod_N2'] = dict(system='euler', Q0=sod(200), tf=0.2, L=[1.], order=2,
                       bts=['transitive'])
    c['sod_short_N3'] = dict(system='euler', Q0=sod(100), tf=0.01, L=[1.], order=3,
                             bts=['transitive'])
    c['euler1d_smooth_N3'] = dict(system='euler', Q0=euler_smooth((64, )), tf=0.02, L=[1.],
                                  order=3, bts=['periodic'])
    c['euler2d_smooth_N3'] = dict(system='euler', Q0=euler_smooth((24, 20)), tf=0.03,
                                  L=[1., 1.], order=3, bts=['periodic', 'periodic'])
    c['euler2d_smooth_N2'] = dict(system='euler', Q0=euler_smooth((24, 20)), tf=0.03,
                                  L=[1., 1.], order=2, bts=['periodic', 'transitive'])
    c['euler2d_explosion_N3'] = dict(system='euler', Q0=euler_explosion((32, 32)), tf=0.03,
                                     L=[1., 1.], order=3, bts=['transitive', 'transitive'])
    c['advect_nc_1d_N3'] = dict(system='advect_nc', Q0=advect_nc_smooth((40, )), tf=0.05,
                                L=[1.], order=3, bts=['periodic'])
    c['advect_nc_2d_N2'] = dict(system='advect_nc', Q0=advect_nc_smooth((16, 12)), tf=0.08,
                                L=[1., 1.], order=2, bts=['periodic', 'periodic'])
    # second-order (viscous) flux F(Q, dQ, d): reference tests/navier_stokes/system.py
    c['ns1d_smooth_N3'] = dict(system='navier_stokes', Q0=ns_smooth((48, )), tf=0.01, L=[1.],
                               order=3, bts=['periodic'], second_order=True)
    c['ns2d_smooth_N2'] = dict(system='navier_stokes', Q0=ns_smooth((16, 12)), tf=0.02,
                               L=[1., 1.], order=2, bts=['periodic', 'transitive'],
                               second_order=True)
    # 3-D: BASELINE config 5 at reduced size (oracle = reference + the zero_index fix)
    c['euler3d_smooth_N2'] = dict(system='euler', Q0=euler_smooth((10, 8, 6)), tf=0.02,
                                  L=[1., 1., 1.], order=2,
                                  bts=['periodic', 'periodic', 'transitive'])
    c['ns3d_taylor_green_N3'] = dict(system='navier_stokes', Q0=taylor_green((6, 6, 6)),
                                     tf=0.05, L=[2 * np.pi] * 3, order=3,
                                     bts=['periodic'] * 3, second_order=True)
    # stiff = True: Newton-Krylov predictor (dg.cpp:173-185)
    c['euler1d_smooth_N3_stiff'] = dict(system='euler', Q0=euler_smooth((32, )), tf=0.02,
                                        L=[1.], order=3, bts=['periodic'], stiff=True)
    # (shifted up: every state has max|Q| >= 1.4, well clear of the reference's
    #  ||x||inf >= 1 Newton termination rule)
    c['advect_nc_2d_N2_stiff'] = dict(system='advect_nc', Q0=advect_nc_smooth((10, 8)) + 0.5,
                                      tf=0.06, L=[1., 1.], order=2,
                                      bts=['periodic', 'periodic'], stiff=True)
    c['reactive1d_smooth_N3_stiff'] = dict(system='reactive_euler',
                                           Q0=reactive_disc((40, ), smooth=True), tf=0.03,
                                           L=[1.], order=3, bts=['transitive'], stiff=True)
    c['reactive2d_disc_N3_stiff'] = dict(system='reactive_euler', Q0=reactive_disc((12, 12)),
                                         tf=0.02, L=[1., 1.], order=3,
                                         bts=['transitive', 'transitive'], stiff=True)
    # Roe and Osher fluxes (fluxes.cpp:20-70)
    c['euler1d_smooth_N3_osher'] = dict(system='euler', Q0=euler_smooth((48, )), tf=0.02, L=[1.],
                                        order=3, bts=['periodic'], flux='osher')
    c['euler2d_smooth_N2_roe'] = dict(system='euler', Q0=euler_smooth((16, 12)), tf=0.03,
                                      L=[1., 1.], order=2, bts=['periodic', 'transitive'],
                                      flux='roe')
    c['sod_short_N2_osher'] = dict(system='euler', Q0=sod(100), tf=0.02, L=[1.], order=2,
                                   bts=['transitive'], flux='osher')
    c['ns1d_smooth_N2_osher'] = dict(system='navier_stokes', Q0=ns_smooth((32, )), tf=0.01,
                                     L=[1.], order=2, bts=['periodic'], second_order=True,
                                     flux='osher')
    # edge cases: scalar system (V = 1), order 1 and 4, tiny grids, a 2-D strip one cell wide
    c['burgers1d_N3'] = dict(system='burgers', Q0=1.5 + smooth_product((40, ))[..., None],
                             tf=0.05, L=[1.], order=3, bts=['periodic'])
    c['burgers2d_N2'] = dict(system='burgers', Q0=1.5 + smooth_product((12, 10))[..., None],
                             tf=0.05, L=[1., 1.], order=2, bts=['periodic', 'transitive'])
    c['euler1d_N1'] = dict(system='euler', Q0=euler_smooth((32, )), tf=0.02, L=[1.], order=1,
                           bts=['periodic'])
    c['euler1d_N4'] = dict(system='euler', Q0=euler_smooth((24, )), tf=0.02, L=[1.], order=4,
                           bts=['transitive'])
    c['euler1d_tiny'] = dict(system='euler', Q0=euler_smooth((3, )), tf=0.1, L=[1.], order=2,
                             bts=['transitive'])
    c['euler2d_strip_N2'] = dict(system='euler', Q0=euler_smooth((40, 1)), tf=0.02, L=[1., 1.],
                                 order=2, bts=['transitive', 'transitive'])
    # BASELINE config 4 at reduced size: GPR model (V = 17, F + B + S), stiff, order 2
    c['gpr1d_N2_stiff'] = dict(system='gpr', Q0=gpr_disc((24, )), tf=0.004, L=[1.], order=2,
                               bts=['transitive'], stiff=True)
    c['gpr2d_N2_stiff'] = dict(system='gpr', Q0=gpr_disc((8, 8)), tf=0.006, L=[1., 1.], order=2,
                               bts=['transitive', 'transitive'], stiff=True)
    # BASELINE config 3 at reduced size: reactive Euler, stiff Newton predictor, Osher flux
    c['reactive2d_disc_N3_stiff_osher'] = dict(system='reactive_euler',
                                               Q0=reactive_disc((10, 10)), tf=0.02,
                                               L=[1., 1.], order=3,
                                               bts=['transitive', 'transitive'], stiff=True,
                                               flux='osher')
    return c


# ---------------------------------------------------------------------------
# BASELINE configs 2-5 at the parity sizes SURVEY 8d names (C2 256^2, C3 64^2, C4 32^2,
# C5 16^3), each run for exactly K = 1, 5 and 10 steps: the final time tf[K] lies in the
# middle of the reference's K-th step (measured with tools/golden_times.py), so both
# sides take K steps, the last one clipped to tf (stepper.cpp:72-73).
# ---------------------------------------------------------------------------
def sized_bases():
    b = {}
    b['c2_explosion_256'] = dict(system='euler', Q0=euler_explosion((256, 256)), L=[1., 1.],
                                 order=3, bts=['transitive', 'transitive'],
                                 tf_guess=26 * 0.2 * 0.9 / (2 * np.sqrt(1.4) * 256),
                                 tf={1: 0.000148563, 5: 0.00119622, 10: 0.00521785})
    b['c3_reactive_64'] = dict(system='reactive_euler', Q0=reactive_disc((64, 64)), L=[1., 1.],
                               order=3, bts=['transitive', 'transitive'], stiff=True,
                               flux='osher', tf_guess=0.02,
                               tf={1: 0.000703125, 5: 0.00615815, 10: 0.0305194})
    b['c4_gpr_32'] = dict(system='gpr', Q0=gpr_disc((32, 32)), L=[1., 1.], order=2,
                          bts=['transitive', 'transitive'], stiff=True, tf_guess=0.02,
                          tf={1: 9.627e-05, 5: 0.000866593, 10: 0.00452912})
    b['c5_taylor_green_16'] = dict(system='navier_stokes', Q0=taylor_green((16, 16, 16)),
                                   L=[2 * np.pi] * 3, order=3, bts=['periodic'] * 3,
                                   second_order=True, tf_guess=0.05,
                                   tf={1: 0.00113089, 5: 0.0101782, 10: 0.0531573})
    return b


def sized_cases():
    """name_K<k> -> case dict (as solver_cases) for K = 1, 5, 10."""
    out = {}
    for name, b in sized_bases().items():
        for K, tf in b['tf'].items():
            c = {k: v for k, v in b.items() if k not in ('tf', 'tf_guess')}
            c['tf'] = tf
            c['steps'] = K
            out['%s_K%d' % (name, K)] = c
    return out
