"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref,
built by `make -C oracle` from /root/reference) — run in the build container:

    python tests/golden/make_golden.py [small] [sized]

The reference ships no golden vectors for this path, so these
fixtures, produced by its own solver with CPU callbacks compiled from
pypde_b200/systems/systems_src.h, pin both the oracle and the CUDA path.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle import reference as R  # noqa: E402


NOISE_SEEDS = (2024, 1, 2, 3)
# nThreads of the reference runs.  Fixed, not the host's core count: the reference's results
# depend on it at rounding level whenever a thread slab starts at an odd row offset (1-D
# N = 3, 64 cells: nThreads 1, 2, 4 agree bit for bit, 3, 5, 7 differ from them by ~4e-13 —
# Eigen's reductions take alignment-dependent paths), which its +-1 ulp self-noise covers.
THREADS = 4


def main():
    out = {}
    # stand-alone WENO (reference api.cpp:32-48)
    out['weno_kat_N2'] = R.weno_solver(cases.weno_kat_input(), 2)
    out['weno_kat_N3'] = R.weno_solver(cases.weno_kat_input(), 3)
    out['weno_rand_1d_N4'] = R.weno_solver(cases.weno_random((15, 2)), 4)
    out['weno_rand_2d_N3'] = R.weno_solver(cases.weno_random((9, 8, 2)), 3)
    out['weno_rand_2d_N2'] = R.weno_solver(cases.weno_random((6, 7, 3)), 2)
    np.savez_compressed(os.path.join(HERE, 'weno.npz'), **out)

    # The reference's own round-off self-noise on a case: the same run with every entry of
    # the initial data moved by +-1 ulp, for NOISE_SEEDS independent
    # perturbations; the maximum is stored next to the fixture.  Parity tolerances are
    # max(stated tolerance, 4 x this)  (tests/conftest.py: parity_tolerance).
    def solve_and_noise(name, c, sol):
        ndim = c['Q0'].ndim - 1
        lib = 'libpypde_ref3d.so' if ndim == 3 else 'libpypde_ref.so'
        F, B, S = R.system_callbacks(c['system'], ndim)

        def run(Q0, count_steps=False):
            def go():
                return R.pde_solver(Q0, c['tf'], c['L'], F=F, B=B, S=S, boundaryTypes=c['bts'],
                                    order=c['order'], ndt=1, flux=c.get('flux', 'rusanov'),
                                    stiff=c.get('stiff', False), nThreads=THREADS,
                                    secondOrder=c.get('second_order', False), lib=lib)[0]
            if not count_steps:
                return go(), None
            # the reference prints "t = ..." per step (iterator.cpp:134): count them
            sys.stdout.flush()
            saved = os.dup(1)
            tmp = tempfile.TemporaryFile('w+b')
            os.dup2(tmp.fileno(), 1)
            try:
                out = go()
            finally:
                os.dup2(saved, 1)
                os.close(saved)
            tmp.seek(0)
            n = sum(1 for l in tmp.read().decode().splitlines() if l.startswith('t = '))
            tmp.close()
            return out, n

        sol[name], steps = run(c['Q0'], count_steps='steps' in c)
        if 'steps' in c:
            assert steps == c['steps'], (name, steps, c['steps'])
            sol[name + '__steps'] = np.array(steps)
        noises = []
        for seed in NOISE_SEEDS:
            rng = np.random.default_rng(seed)
            Qp = np.where(rng.random(c['Q0'].shape) < 0.5, np.nextafter(c['Q0'], np.inf),
                          np.nextafter(c['Q0'], -np.inf))
            noises.append(float(np.abs(run(Qp)[0] - sol[name]).max() / np.abs(sol[name]).max()))
        sol[name + '__noise'] = np.array(max(noises))
        sol[name + '__noises'] = np.array(noises)
        print('%-28s %-14s max|u| %.6f  self-noise max %.2e  (%s)' %
              (name, sol[name].shape[:-1], float(np.abs(sol[name]).max()), max(noises),
               ' '.join('%.1e' % x for x in noises)), flush=True)

    which = sys.argv[1:] or ['small', 'sized']
    if 'small' in which:
        sol = {}
        for name, c in cases.solver_cases().items():
            solve_and_noise(name, c, sol)
        np.savez_compressed(os.path.join(HERE, 'solver.npz'), **sol)
    if 'sized' in which:
        # BASELINE configs 2-5 at SURVEY 8d's parity sizes, K = 1, 5, 10 steps
        sol = {}
        for name, c in cases.sized_cases().items():
            solve_and_noise(name, c, sol)
        np.savez_compressed(os.path.join(HERE, 'solver_sized.npz'), **sol)

    # tables of the reference (poly/basis.cpp etc.) for N = 2, 3, 4
    st = R.Stages()
    tab = {}
    for N in (2, 3, 4):
        for k, v in st.tables(N).items():
            tab['N%d_%s' % (N, k)] = v
    np.savez_compressed(os.path.join(HERE, 'tables.npz'), **tab)


if __name__ == '__main__':
    main()
