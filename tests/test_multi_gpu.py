"""Slab decomposition across ranks.

CPU part (gloo, world_size 2, no GPU): the host mirror of the decomposition
(pypde_b200/slabs.py — same neighbour, halo-width and ghost-source rules as
Solver::exchange_halos / k_boundaries) exchanges real halos between two
processes; every rank advances its slab with the oracle; the stitched result
must equal the undivided oracle step bit for bit (halo width N is
necessary and sufficient, dt is the max over ranks).

GPU part (marked gpu, needs >= 2 devices): the same check on the CUDA path with
NCCL, through tools/multi_gpu_check.py under torchrun.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
from oracle import ader_weno as O
from oracle import systems as SY
from pypde_b200 import slabs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_cover_the_axis():
    for n0 in (7, 64, 2048):
        for world in (1, 2, 3, 8):
            b = [slabs.slab_bounds(n0, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n0
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))


def test_halo_modes_and_neighbours():
    assert slabs.halo_modes(0, 1, True) == (slabs.WRAP, slabs.WRAP)
    assert slabs.halo_modes(0, 1, False) == (slabs.CLAMP, slabs.CLAMP)
    assert slabs.halo_modes(0, 4, False) == (slabs.CLAMP, slabs.HALO)
    assert slabs.halo_modes(3, 4, False) == (slabs.HALO, slabs.CLAMP)
    assert slabs.halo_modes(0, 4, True) == (slabs.HALO, slabs.HALO)
    assert slabs.neighbours(0, 4, True) == (3, 1)
    assert slabs.neighbours(3, 4, True) == (2, 0)
    assert slabs.neighbours(0, 4, False) == (None, 1)


WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests', 'golden'))
import cases
from oracle import ader_weno as O, systems as SY
from pypde_b200 import slabs

dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
periodic = %(periodic)r
N = %(N)d
shape = %(shape)r
s = SY.SYSTEMS['euler'](2)
u_full = cases.euler_smooth(shape)
dX = np.array([1. / n for n in shape])
bt = [1 if periodic else 0, 0]
u = slabs.split(u_full, world)[rank]
t = 0.
for k in range(3):
    hlo, hhi = slabs.exchange_halos_torch(u, N, periodic)
    modes = slabs.halo_modes(rank, world, periodic)
    ub0 = slabs.padded_axis0(u, hlo, hhi, N, modes)
    # transverse ghosts as boundaries.cpp, axis 0 already padded
    ub = np.pad(ub0, [(0, 0), (N, N), (0, 0)], mode='edge')
    w = O.weno(ub, N, 2)
    import torch
    mx = torch.tensor([O.cfl_max(w, s['F'], None, dX, N, False)], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dt = 0.9 / float(mx.item())
    if k <= 5:
        dt *= 0.2
    qh = O.predictor(w, dt, s['F'], None, None, dX, N).reshape(w.shape[:2] + (N, N, N, 4))
    u = O.fv_apply(u, qh, dt, s['F'], None, None, dX, N)
    t += dt
np.save(%(out)r %% rank, u)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize('periodic', [False, True])
def test_two_rank_slabs_reproduce_undivided_step_gloo(tmp_path, periodic):
    N, shape = 3, (16, 10)
    out = str(tmp_path / 'u%d.npy')
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % dict(root=ROOT, periodic=periodic, N=N, shape=shape, out=out))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', OMP_NUM_THREADS='1')
    port = 29500 + (os.getpid() % 500) + (7 if periodic else 0)
    subprocess.check_call([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                           '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                           '--master-port', str(port), str(script)], env=env, timeout=600)
    parts = [np.load(out % r) for r in range(2)]
    # undivided oracle
    s = SY.SYSTEMS['euler'](2)
    u = cases.euler_smooth(shape)
    dX = np.array([1. / n for n in shape])
    t = 0.
    for k in range(3):
        u, dt = O.step(u, t, k, 1e9, dX, [1 if periodic else 0, 0], s['F'], None, None, N, 0.9)
        t += dt
    assert np.array_equal(slabs.stitch(parts), u)


@pytest.mark.gpu
def test_two_gpu_slabs_match_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (covered on CPU by the gloo test; run with gpurun --gpus 2)')
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    out = subprocess.check_output([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                                   '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                                   '--master-port', '29611',
                                   os.path.join(ROOT, 'tools', 'multi_gpu_check.py')], env=env,
                                  timeout=900).decode()
    assert 'MULTI_GPU_CHECK OK' in out, out
