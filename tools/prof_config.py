"""A few steps of one bench.py configuration with the state resident in HBM — the target of
the ncu captures (profiles/README.md):

    ncu --set full --clock-control none --import-source on -k regex:k_dg_stiff -s 1 -c 1 \
        -o gpurun_out/prof_c3 python tools/prof_config.py c3 [steps] [size]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pypde_b200.handle import Solver  # noqa: E402
from pypde_b200.systems import cuda_sources  # noqa: E402

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
size = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = bench.CONFIGS[name]
os.environ['PYPDE_B200_QUIET'] = '1'
F, B, S, V = cuda_sources(cfg['system'], len(cfg['shape']))
gshape, rows, Q0, dX = bench.slab_problem(cfg, 0, 1, size)
sol = Solver(Q0.shape, None, F=F, B=B, S=S, boundaryTypes=cfg['bts'], cfl=0.9, order=cfg['order'],
             dX=dX, flux=cfg['flux'], stiff=cfg['stiff'])
sol.set_state(Q0)
sol.begin(1e9)
for _ in range(steps):
    sol.step_async()
print(name, Q0.shape, sol.sync())
