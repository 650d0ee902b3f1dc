"""One k_wavespeeds-heavy step of C5 (3-D Navier-Stokes, 48^3) for ncu."""
import sys
import numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests/golden')
import cases
from pypde_b200.handle import Solver
from pypde_b200.systems import cuda_sources
n = 48
F, B, S, V = cuda_sources('navier_stokes', 3)
Q0 = cases.taylor_green((n, n, n))
sol = Solver(Q0.shape, [2 * np.pi] * 3, F=F, boundaryTypes=['periodic'] * 3, order=3)
sol.set_state(Q0)
sol.begin(1e9)
for _ in range(3):
    sol.step_async()
print(sol.sync())
