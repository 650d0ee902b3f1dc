"""Throughput survey of the five BASELINE configs at single-GPU sizes (a few
steps each, device-resident state): cell-updates/s and per-kernel ms."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
sys.path.insert(0, 'tests/golden')
import cases
from pypde_b200.handle import Solver
from pypde_b200.systems import cuda_sources


def run(name, system, Q0, L, N, bts, steps=4, warm=2, **kw):
    ndim = Q0.ndim - 1
    F, B, S, V = cuda_sources(system, ndim)
    t0 = time.time()
    sol = Solver(Q0.shape, L, F=F, B=B, S=S, boundaryTypes=bts, order=N, **kw)
    tjit = time.time() - t0
    sol.set_state(Q0)
    sol.begin(1e9)
    for _ in range(warm):
        sol.step_async()
    sol.sync()
    sol.set_profiling(True)
    t0 = time.time()
    for _ in range(steps):
        sol.step_async()
    t, dt, nan = sol.sync()
    wall = time.time() - t0
    kt = sol.kernel_times()
    ms = sum(v[0] for v in kt.values()) / steps
    cells = int(np.prod(Q0.shape[:-1]))
    print('%-34s %-12s create %.1fs  %.2f ms/step (wall %.2f)  %.3e cu/s  nan=%s' %
          (name, 'x'.join(map(str, Q0.shape[:-1])), tjit, ms, wall / steps * 1e3,
           cells / (ms * 1e-3), nan))
    print('      ' + '  '.join('%s=%.2f' % (k.replace('k_', ''), v[0] / steps)
                               for k, v in kt.items() if v[0] / steps > 0.05))
    sol.close()


if __name__ == '__main__':
    big = len(sys.argv) > 1 and sys.argv[1] == 'big'
    only = sys.argv[2].split(',') if len(sys.argv) > 2 else None   # e.g. C4,C5
    if only:
        _run = run

        def run(name, *a, **kw):
            if name[:2] in only:
                _run(name, *a, **kw)
    n2 = 1024 if big else 256
    run('C1 1-D Euler Sod N=2', 'euler', cases.sod(200), [1.], 2, ['transitive'], steps=20)
    run('C2 2-D Euler explosion N=3', 'euler', cases.euler_explosion((n2, n2)), [1., 1.], 3,
        ['transitive'] * 2)
    run('C2 smooth periodic variant', 'euler', cases.euler_smooth((n2, n2)), [1., 1.], 3,
        ['periodic'] * 2)
    n3 = 512 if big else 128
    run('C3 reactive Euler stiff Osher N=3', 'reactive_euler', cases.reactive_disc((n3, n3)),
        [1., 1.], 3, ['transitive'] * 2, stiff=True, flux='osher')
    run('C3 (same, Rusanov)', 'reactive_euler', cases.reactive_disc((n3, n3)), [1., 1.], 3,
        ['transitive'] * 2, stiff=True)
    n4 = 256 if big else 64
    run('C4 GPR stiff N=2', 'gpr', cases.gpr_disc((n4, n4), smooth=False), [1., 1.], 2,
        ['transitive'] * 2, stiff=True)
    n5 = 64 if big else 24
    run('C5 3-D Navier-Stokes N=3', 'navier_stokes', cases.taylor_green((n5, n5, n5)),
        [2 * np.pi] * 3, 3, ['periodic'] * 3)
