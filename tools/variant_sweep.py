"""Kernel variants (selected by the library's environment switches) on one bench.py
configuration: per-kernel ms per step, whether the result has the bits of the first
variant, and — for the stiff set with PYPDE_B200_STIFF_STATS=1 — iteration counters per cell.

    python tools/variant_sweep.py stiff  [config=c3] [size=512] [steps=4]
    python tools/variant_sweep.py eig    [config=c4] [size=256] [steps=4]
    python tools/variant_sweep.py faces  [config=c2] [size=2048] [steps=4]
    python tools/variant_sweep.py weno3d [config=c5] [size=128] [steps=4]   (32 x size x size)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from pypde_b200.handle import Solver  # noqa: E402
from pypde_b200.systems import cuda_sources  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'stiff'
name = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != 'x' else {'stiff': 'c3', 'eig': 'c4', 'faces': 'c2', 'weno3d': 'c5', 'c5': 'c5', 'occ': 'c4', 'gprstiff': 'c4', 'eig2': 'c4', 'osher': 'c3', 'stiff2': 'c3'}[which]
size = int(sys.argv[3]) if len(sys.argv) > 3 else {'stiff': 512, 'eig': 256, 'faces': 2048, 'weno3d': 128, 'c5': 128, 'occ': 256, 'gprstiff': 512, 'eig2': 512, 'osher': 512, 'stiff2': 512}[which]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
cfg = bench.CONFIGS[name]
os.environ['PYPDE_B200_QUIET'] = '1'
F, B, S, V = cuda_sources(cfg['system'], len(cfg['shape']))
gshape, rows, Q0, dX = bench.slab_problem(cfg, 0, 1, size)
cells = int(np.prod(Q0.shape[:-1]))

SETS = {}
SETS['eig'] = [
    ('two-pass Jacobian, active block stored compact (default)', {}),
    ('one pass: full matrix in local memory, then gathered', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_EIG_TWOPASS=0'}),
    ('  permutation step by exchanges in memory instead of bit masks', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_EIG_MASK=0'}),
    ('  QR iteration instead of the certified characteristic polynomial', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_EIG_HESS_POLY=0'}),
    ('round 1: QR iteration on the full matrix', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_EIG_DEFLATE=0'}),
    ('default, k_wavespeeds 128 x 4 blocks per SM', {'PYPDE_B200_WS_BLOCK': '128', 'PYPDE_B200_WS_MINBLOCKS': '4'}),
    ('default, 128 x 3 blocks per SM (smem pad)', {'PYPDE_B200_WS_BLOCK': '128', 'PYPDE_B200_WS_MINBLOCKS': '4', 'PYPDE_B200_WS_SMEM_PAD': '73000'}),
    ('default, 128 x 2 blocks per SM (smem pad)', {'PYPDE_B200_WS_BLOCK': '128', 'PYPDE_B200_WS_MINBLOCKS': '4', 'PYPDE_B200_WS_SMEM_PAD': '110000'}),
    ('default, 128 x 1 block per SM (smem pad)', {'PYPDE_B200_WS_BLOCK': '128', 'PYPDE_B200_WS_MINBLOCKS': '4', 'PYPDE_B200_WS_SMEM_PAD': '200000'}),
    ('default, 256 x 1 block per SM (smem pad)', {'PYPDE_B200_WS_BLOCK': '256', 'PYPDE_B200_WS_MINBLOCKS': '2', 'PYPDE_B200_WS_SMEM_PAD': '120000'}),
]
SETS['eig2'] = [   # details of the n > 5 path
    ('default', {}),
    ('pass 1 of the Jacobian as a loop (one inlined flux instead of V)', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_EIG_PASS1_ROLLED=1'}),
    ('active block with row pitch m instead of V', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_EIG_PITCH_FULL=0'}),
    ('Hessenberg reduction with zero-factor skips', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_EIG_HESS_SKIP=1'}),
    ('at most 3 balancing sweeps instead of 6', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_EIG_BAL_SWEEPS=3'}),
    ('at most 1 balancing sweep', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_EIG_BAL_SWEEPS=1'}),
]
SETS['osher'] = [   # Osher / Roe dissipation |A| x for V = 3..5
    ('projector form where it certifies (default)', {}),
    ('real-Schur route on every matrix (elmhes / hqr2 / QR solve)', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_ABS_POLY=0'}),
    ('default, k_faces 4 blocks per SM', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_FACES_MINBLOCKS=4'}),
    ('default, k_faces 2 blocks per SM', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_FACES_MINBLOCKS=2'}),
]
SETS['stiff2'] = [   # the Gram-Schmidt loop of k_dg_stiff
    ('next basis vector prefetched during the reduction (default)', {}),
    ('no prefetch', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_STIFF_PREFETCH=0'}),
    ('quotients instead of reciprocal multiplies in the Krylov loop', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_STIFF_RECIP=0'}),
    ('the direction z re-read from the basis instead of kept in registers', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_STIFF_ZREG=0'}),
    ('DMUL + DADD pairs instead of fused multiply-adds in the Krylov linear algebra', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_STIFF_FMA=0'}),
    ('all four off (the kernel of session S)', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_STIFF_ZREG=0;PDE_STIFF_RECIP=0;PDE_STIFF_PREFETCH=0;PDE_STIFF_FMA=0'}),
]
SETS['occ'] = [   # occupancy against registers for the n > 5 wave-speed kernel (latency-bound per thread)
    ('default: k_wavespeeds 512 x 1 (128 registers)', {}),
    ('640 x 1 (96 registers)', {'PYPDE_B200_WS_BLOCK': '640', 'PYPDE_B200_WS_MINBLOCKS': '1'}),
    ('768 x 1 (80 registers)', {'PYPDE_B200_WS_BLOCK': '768', 'PYPDE_B200_WS_MINBLOCKS': '1'}),
    ('256 x 3 (85 registers)', {'PYPDE_B200_WS_BLOCK': '256', 'PYPDE_B200_WS_MINBLOCKS': '3'}),
    ('1024 x 1 (64 registers)', {'PYPDE_B200_WS_BLOCK': '1024', 'PYPDE_B200_WS_MINBLOCKS': '1'}),
    ('512 x 2 (64 registers)', {'PYPDE_B200_WS_BLOCK': '512', 'PYPDE_B200_WS_MINBLOCKS': '2'}),
    ('256 x 4 (64 registers)', {'PYPDE_B200_WS_BLOCK': '256', 'PYPDE_B200_WS_MINBLOCKS': '4'}),
]
SETS['gprstiff'] = [
    ('default (254 registers, 2 blocks of 4 warps per SM)', {}),
    ('3 blocks per SM (170 registers)', {'PYPDE_B200_STIFF_MINBLOCKS': '3'}),
    ('4 blocks per SM (128 registers)', {'PYPDE_B200_STIFF_MINBLOCKS': '4'}),
    ('8 warps per block, 1 block', {'PYPDE_B200_STIFF_WPB': '8', 'PYPDE_B200_STIFF_MINBLOCKS': '1'}),
]
SETS['c5'] = [
    ('k_faces_side + k_dg_g (default)', {}),
    ('k_wavespeeds + k_faces', {'PYPDE_B200_FUSED_FACES': '0'}),
    ('k_dg (space-time node per thread)', {'PYPDE_B200_DG_NODE': '0'}),
    ('k_dg_g 12 blocks per SM (170 registers)', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_DGG_MINBLOCKS=12'}),
    ('k_dg_g 16 blocks per SM (128 registers)', {'PYPDE_B200_EXTRA_DEFINES': 'PDE_DGG_MINBLOCKS=16'}),
    ('k_faces_side 256 x 2', {'PYPDE_B200_FS_BLOCK': '256', 'PYPDE_B200_FS_MINBLOCKS': '2'}),
    ('k_faces_side 1024 x 1 (64 registers)', {'PYPDE_B200_FS_BLOCK': '1024', 'PYPDE_B200_FS_MINBLOCKS': '1'}),
]
SETS['faces'] = [
    ('round 1 equivalent: k_cfl on w', {'PYPDE_B200_CFL_Q': '0'}),
    ('k_cfl_q on k_weno2d\'s cell averages (default)', {}),
    ('  fs_block 128 x 4', {'PYPDE_B200_FS_BLOCK': '128', 'PYPDE_B200_FS_MINBLOCKS': '4'}),
    ('  fs_block 512 x 1', {'PYPDE_B200_FS_BLOCK': '512', 'PYPDE_B200_FS_MINBLOCKS': '1'}),
    ('k_faces_fused', {'PYPDE_B200_FACES_SIDE': '0'}),
]
SETS['weno3d'] = [
    ('three k_weno_sweep launches (default)', {}),
    ('k_weno3d tile 4x4x8', {'PYPDE_B200_WENO3D': '1'}),
    ('k_weno3d tile 2x4x8', {'PYPDE_B200_WENO3D': '1', 'PYPDE_B200_W3_TILE': '2,4,8'}),
    ('k_weno3d tile 4x4x4', {'PYPDE_B200_WENO3D': '1', 'PYPDE_B200_W3_TILE': '4,4,4'}),
    ('k_weno3d tile 2x2x8', {'PYPDE_B200_WENO3D': '1', 'PYPDE_B200_W3_TILE': '2,2,8'}),
]
SETS['stiff'] = [
    ('default (1 resident Krylov vector, 4 x 4 warps, 128 registers)', {}),
    ('+ stats', {'PYPDE_B200_STIFF_STATS': '1'}),
    ('KS=2', {'PYPDE_B200_STIFF_KS': '2'}),
    ('KS=3', {'PYPDE_B200_STIFF_KS': '3'}),
    ('WPB=8 minblocks=2', {'PYPDE_B200_STIFF_WPB': '8', 'PYPDE_B200_STIFF_MINBLOCKS': '2'}),
    ('WPB=2 minblocks=8', {'PYPDE_B200_STIFF_WPB': '2', 'PYPDE_B200_STIFF_MINBLOCKS': '8'}),
]
ref = None
for label, env in SETS[which]:
    os.environ.update(env)
    try:
        sol = Solver(Q0.shape, None, F=F, B=B, S=S, boundaryTypes=cfg['bts'], cfl=0.9,
                     order=cfg['order'], dX=dX, flux=cfg['flux'], stiff=cfg['stiff'])
    except RuntimeError as ex:
        print('%-44s failed: %s' % (label, str(ex).splitlines()[0]))
        continue
    finally:
        for k in env:
            del os.environ[k]
    sol.set_state(Q0)
    sol.begin(1e9)
    for _ in range(2):
        sol.step_async()
    sol.sync()
    sol.set_profiling(True)
    for _ in range(steps):
        sol.step_async()
    t, dt, nan = sol.sync()
    kt = sol.kernel_times()
    u = sol.get_state()
    if ref is None:
        ref = u
    line = '%-44s step %8.3f ms  same bits as the first: %-5s  %s' % (
        label, sum(v[0] for v in kt.values()) / steps, np.array_equal(u, ref),
        '  '.join('%s=%.3f' % (k.replace('k_', ''), v[0] / steps) for k, v in kt.items()
                  if v[0] / steps > 0.02))
    if env.get('PYPDE_B200_STIFF_STATS') == '1':
        st = sol.stiff_stats()
        ncw = int(np.prod([n + 2 for n in Q0.shape[:-1]])) * (steps + 2)
        line += '  per cell: newton %.2f obj %.2f inner %.2f beyond-smem %.4f' % tuple(
            st[k] / ncw for k in ('newton', 'obj', 'inner', 'deep'))
    print(line, flush=True)
    sol.close()
