"""Launch-bound regime: wall time per step of C1 (1-D Euler Sod, 200 cells, order 2) and of a
64 x 64 2-D Euler grid, device-resident state, no per-kernel events.  Run with
PYPDE_B200_GRAPH=1 / 0 to compare the CUDA-graph replay of the step with plain launches."""
import os
import sys
import time

sys.path.insert(0, '.')
sys.path.insert(0, 'tests/golden')
import cases
from pypde_b200.handle import Solver
from pypde_b200.systems import cuda_sources

for name, Q0, L, N in [('C1 sod 200 N=2', cases.sod(200), [1.], 2),
                       ('2-D Euler 64x64 N=3', cases.euler_smooth((64, 64)), [1., 1.], 3)]:
    ndim = Q0.ndim - 1
    F, B, S, V = cuda_sources('euler', ndim)
    sol = Solver(Q0.shape, L, F=F, boundaryTypes=['transitive'] * ndim, order=N)
    sol.set_state(Q0)
    sol.begin(1e9)
    for _ in range(50):
        sol.step_async()
    sol.sync()
    K = 2000
    t0 = time.perf_counter()
    for _ in range(K):
        sol.step_async()
    sol.sync()
    dt = time.perf_counter() - t0
    print('%-22s PYPDE_B200_GRAPH=%s  %.1f us/step  %.3e cell-updates/s' %
          (name, os.environ.get('PYPDE_B200_GRAPH', 'default'), dt / K * 1e6,
           Q0.size / V * K / dt))
    sol.close()
