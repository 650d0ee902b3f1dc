"""A few steps of C3 (2-D reactive Euler, stiff Newton-Krylov predictor, 192^2) for ncu."""
import sys
sys.path.insert(0, '.')
sys.path.insert(0, 'tests/golden')
import cases
from pypde_b200.handle import Solver
from pypde_b200.systems import cuda_sources
n = 192
F, B, S, V = cuda_sources('reactive_euler', 2)
Q0 = cases.reactive_disc((n, n))
sol = Solver(Q0.shape, [1., 1.], F=F, B=B, S=S, boundaryTypes=['transitive'] * 2, order=3, stiff=True)
sol.set_state(Q0)
sol.begin(1e9)
for _ in range(3):
    sol.step_async()
print(sol.sync())
