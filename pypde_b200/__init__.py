"""pypde_b200 — B200-native drop-in for the ADER-WENO path of PyPDE.

Public API = the reference's (`pypde/__init__.py:1`): `pde_solver`,
`weno_solver`; plus the handle-based `Solver` for state resident in HBM.
"""
from pypde_b200.solvers import pde_solver, weno_solver  # noqa: F401
from pypde_b200.cfuncs import CudaSource, DeviceFunction  # noqa: F401
