"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref,
built by `make -C oracle` from /root/reference) — run in the build container:

    python tests/golden/make_golden.py

The reference ships no golden vectors for this path, so these
fixtures, produced by its own solver with CPU callbacks compiled from
pypde_b200/systems/systems_src.h, pin both the oracle and the CUDA path.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle import reference as R  # noqa: E402


def main():
    out = {}
    # stand-alone WENO (reference api.cpp:32-48)
    out['weno_kat_N2'] = R.weno_solver(cases.weno_kat_input(), 2)
    out['weno_kat_N3'] = R.weno_solver(cases.weno_kat_input(), 3)
    out['weno_rand_1d_N4'] = R.weno_solver(cases.weno_random((15, 2)), 4)
    out['weno_rand_2d_N3'] = R.weno_solver(cases.weno_random((9, 8, 2)), 3)
    out['weno_rand_2d_N2'] = R.weno_solver(cases.weno_random((6, 7, 3)), 2)
    np.savez_compressed(os.path.join(HERE, 'weno.npz'), **out)

    sol = {}
    rng = np.random.default_rng(2024)
    for name, c in cases.solver_cases().items():
        ndim = c['Q0'].ndim - 1
        lib = 'libpypde_ref3d.so' if ndim == 3 else 'libpypde_ref.so'
        F, B, S = R.system_callbacks(c['system'], ndim)

        def run(Q0):
            return R.pde_solver(Q0, c['tf'], c['L'], F=F, B=B, S=S, boundaryTypes=c['bts'],
                                order=c['order'], ndt=1, flux=c.get('flux', 'rusanov'),
                                stiff=c.get('stiff', False), nThreads=4,
                                secondOrder=c.get('second_order', False), lib=lib)[0]

        sol[name] = run(c['Q0'])
        # the reference's own round-off self-noise: the same run with every entry of
        # the initial data moved by +-1 ulp.  Parity tolerances are
        # max(stated tolerance, 4 x this).
        Qp = np.where(rng.random(c['Q0'].shape) < 0.5, np.nextafter(c['Q0'], np.inf),
                      np.nextafter(c['Q0'], -np.inf))
        noise = float(np.abs(run(Qp) - sol[name]).max() / np.abs(sol[name]).max())
        sol[name + '__noise'] = np.array(noise)
        print('%-26s %-14s max|u| %.6f  self-noise %.2e' % (name, sol[name].shape[:-1],
                                                           float(np.abs(sol[name]).max()), noise))
    np.savez_compressed(os.path.join(HERE, 'solver.npz'), **sol)

    # tables of the reference (poly/basis.cpp etc.) for N = 2, 3, 4
    st = R.Stages()
    tab = {}
    for N in (2, 3, 4):
        for k, v in st.tables(N).items():
            tab['N%d_%s' % (N, k)] = v
    np.savez_compressed(os.path.join(HERE, 'tables.npz'), **tab)


if __name__ == '__main__':
    main()
