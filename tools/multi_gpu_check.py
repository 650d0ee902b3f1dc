"""Run under torchrun with W ranks: each rank advances its slab with the CUDA
path (NCCL halo exchange + dt all-reduce inside libpypde.so) through the
pde_solver C ABI; rank 0 compares the stitched result with its own undivided
single-GPU run.  Prints MULTI_GPU_CHECK OK when bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import cases
import pypde_b200
from pypde_b200 import slabs
from pypde_b200.handle import Solver, comm_init_from_torch, _lib
from pypde_b200.systems import cuda_sources

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
torch.zeros(1, device='cuda')
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
os.environ['PYPDE_B200_QUIET'] = '1'
ok = True
for (shape, N, periodic) in [((64, 48), 3, False), ((64, 48), 3, True), ((32, 40), 2, True)]:
    F, B, S, V = cuda_sources('euler', 2)
    Q0 = cases.euler_smooth(shape)
    bts = ['periodic' if periodic else 'transitive', 'transitive']
    tf = 0.02
    # undivided run on rank 0's GPU (no communicator yet)
    full = None
    if rank == 0:
        full = pypde_b200.pde_solver(Q0.copy(), tf, [1., 1.], F=F, boundaryTypes=bts, order=N,
                                     ndt=1, stiff=False)[0]
    dist.barrier()
    comm_init_from_torch()
    mine = slabs.split(Q0, world)[rank]
    # the slab keeps the global cell size: L_local = rows_local * dx
    Lloc = [mine.shape[0] / shape[0], 1.]
    out = pypde_b200.pde_solver(mine, tf, Lloc, F=F, boundaryTypes=bts, order=N, ndt=1,
                                stiff=False)[0]
    _lib().pypde_b200_comm_finalize()
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        st = slabs.stitch(gathered)
        same = np.array_equal(st, full)
        err = np.abs(st - full).max()
        print('shape', shape, 'N', N, 'periodic', periodic, 'world', world, 'bit-identical', same,
              'max abs diff %.3e' % err)
        ok = ok and same
if rank == 0:
    print('MULTI_GPU_CHECK OK' if ok else 'MULTI_GPU_CHECK FAILED')
dist.barrier()
dist.destroy_process_group()
