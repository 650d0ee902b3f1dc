#!/usr/bin/env python
"""bench.py — cell-updates/s of the ADER-WENO step (BASELINE.json metric).

    python bench.py [--config c2] [--gpus N] [--steps K] [--warmup W]      (our CUDA path)
    python bench.py --impl reference [--config c2] [--steps K] [--warmup W] (reference CPU path)

--config selects one of BASELINE.json's configurations (default c2, the one the metric is
quoted on; the others are reported under profiles/):

    c1        1-D Euler Sod shock tube, 200 cells, order 2, Rusanov
    c2        2-D Euler cylindrical explosion, 2048^2 cells per GPU, order 3, Rusanov
    c2smooth  the same on the smooth periodic data of SURVEY 8d (two Picard iterations in
              every cell, no plateaus)
    c3        2-D reactive Euler, 1024^2, order 3, stiff Newton predictor, Osher flux
    c4        2-D GPR model (V = 17, B, stiff sources), 512^2, order 2 — the SAME 512^2 grid on
              1/2/4/8 GPUs (strong scaling)
    c5        3-D Navier-Stokes (second-order flux), Taylor-Green, order 3, slabs of 32 planes
              of 256^2 per GPU: 256^3 on 8 GPUs (weak scaling)

With N > 1 ranks (one process per GPU under torchrun) the domain is slab-partitioned along
axis 0: every step does a real NCCL halo exchange of N rows of u per side and max-all-reduces
for dt.  Before the timed region a multi-rank run also checks, on small grids, that slab
runs reproduce the undivided run bit for bit (`slab_bit_identical` in the JSON line).

A step is one pass of the hot path (ghost cells, WENO, CFL/dt, DG predictor, interface
fluxes, update).  `value` is measured with the state resident in HBM; `e2e` goes through the
reference-facing C ABI (`pde_solver`) with host buffers.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import cases  # noqa: E402


# ---------------------------------------------------------------------------
# the configurations (BASELINE.json configs[0..4]; SURVEY 8d for the initial data)
# ---------------------------------------------------------------------------
def _ref_tf_explosion(n, k):
    # explosion IC at rest: lambda_max = c = sqrt(1.4), dx = 1/n; the first 6 steps use
    # 0.2 dt (stepper.cpp:69-70)
    dt0 = 0.2 * 0.9 / (2 * np.sqrt(1.4) * n)
    return (min(k, 6) + 5 * max(0, k - 6)) * dt0 * 0.999


def _ref_tf_smooth(n, k):
    # smooth IC: |u| + c = 1 + 1.18, |v| + c = 0.5 + 1.18 at rho = 1
    dt0 = 0.2 * 0.9 / ((2.19 + 1.69) * n)
    return (min(k, 6) + 5 * max(0, k - 6)) * dt0 * 0.999


CONFIGS = {
    'c1': dict(
        workload='1-D Euler Sod shock tube, 200 cells, order 2, Rusanov, FP64, transmissive '
                 '(BASELINE configs[0]); every rank runs its own tube for N>1',
        system='euler', ic=lambda shape, rows=None: cases.sod(shape[0]), shape=(200, ), L=[1.],
        order=2, flux='rusanov', stiff=False, bts=['transitive'], scaling='weak', mode='replica',
        ref=dict(shape=(200, ), tf=lambda k: 0.2, what='the whole case (tf = 0.2, 102 steps)')),
    'c2': dict(
        workload='2-D Euler cylindrical explosion, 2048^2 cells per GPU, order 3, Rusanov, FP64, '
                 'transmissive (BASELINE configs[1]); slab-partitioned along axis 0 for N>1',
        system='euler', ic=cases.euler_explosion, shape=(2048, 2048), L=[1., 1.], order=3,
        flux='rusanov', stiff=False, bts=['transitive', 'transitive'], scaling='weak',
        mode='stack',
        ref=dict(shape=(128, 128), tf=lambda k: _ref_tf_explosion(128, k),
                 what='128x128 cells of the same workload')),
    'c2smooth': dict(
        workload='2-D Euler, smooth periodic data rho = 1 + 0.2 sin 2 pi x sin 2 pi y (SURVEY 8d '
                 'variant of BASELINE configs[1]), 2048^2 cells per GPU, order 3, Rusanov, FP64',
        system='euler', ic=cases.euler_smooth, shape=(2048, 2048), L=[1., 1.], order=3,
        flux='rusanov', stiff=False, bts=['periodic', 'periodic'], scaling='weak', mode='stack',
        ref=dict(shape=(128, 128), tf=lambda k: _ref_tf_smooth(128, k),
                 what='128x128 cells of the same workload')),
    'c3': dict(
        workload='2-D reactive Euler (Arrhenius source, K0 = 250, Ea = 2), burnt disc in unburnt '
                 'gas, 1024^2 cells per GPU, order 3, stiff Newton-Krylov predictor, Osher flux, '
                 'FP64, transmissive (BASELINE configs[2])',
        system='reactive_euler', ic=cases.reactive_disc, shape=(1024, 1024), L=[1., 1.], order=3,
        flux='osher', stiff=True, bts=['transitive', 'transitive'], scaling='weak', mode='stack',
        ref=dict(sized='c3_reactive_64', what='64x64 cells of the same workload')),
    'c4': dict(
        workload='2-D GPR continuum model (V = 17, non-conservative B, stiff relaxation sources), '
                 'disc (rho, p) = (4, 4/g) in (2, 2/g), 512^2 cells IN TOTAL, order 2, stiff '
                 'predictor, Rusanov, FP64, transmissive (BASELINE configs[3]); the same grid '
                 'slab-partitioned over N GPUs',
        system='gpr', ic=cases.gpr_disc, shape=(512, 512), L=[1., 1.], order=2, flux='rusanov',
        stiff=True, bts=['transitive', 'transitive'], scaling='strong', mode='split',
        ref=dict(sized='c4_gpr_32', what='32x32 cells of the same workload')),
    'c5': dict(
        workload='3-D compressible Navier-Stokes (second-order viscous flux, mu = 1e-2), '
                 'Taylor-Green vortex on [0, 2 pi]^3, order 3, Rusanov, FP64, periodic; slabs of '
                 '32 planes of 256^2 cells per GPU along axis 0 — 256^3 on 8 GPUs '
                 '(BASELINE configs[4])',
        system='navier_stokes', ic=cases.taylor_green, shape=(32, 256, 256), L=[2 * np.pi] * 3,
        order=3, flux='rusanov', stiff=False, bts=['periodic'] * 3, scaling='weak', mode='grow',
        ref=dict(sized='c5_taylor_green_16', what='16x16x16 cells of the same workload')),
}


def slab_problem(cfg, rank, world, size=None):
    """(global shape, this rank's rows (r0, r1), this rank's initial slab, dX)."""
    shape = tuple(cfg['shape'])
    if size:
        shape = (size, ) * len(shape) if cfg['mode'] != 'grow' else (shape[0], size, size)
    n0 = shape[0]
    if cfg['mode'] in ('stack', 'replica'):
        # every rank holds the same n0-row problem; under 'stack' the slabs are joined along
        # axis 0 (N explosions one above the other), under 'replica' they stay independent
        Q0 = cfg['ic'](shape)
        gshape = (n0 * world, ) + shape[1:] if cfg['mode'] == 'stack' else shape
        dX = [cfg['L'][i] / shape[i] for i in range(len(shape))]
        return gshape, (rank * n0, (rank + 1) * n0), Q0, dX
    if cfg['mode'] == 'split':
        if n0 % world:
            raise SystemExit('bench.py: %d rows do not split over %d ranks' % (n0, world))
        rows = (rank * n0 // world, (rank + 1) * n0 // world)
        gshape = shape
    else:  # 'grow': n0 rows per rank of a domain of fixed physical size
        gshape = (n0 * world, ) + shape[1:]
        rows = (rank * n0, (rank + 1) * n0)
    Q0 = cfg['ic'](gshape, rows=rows)
    dX = [cfg['L'][i] / gshape[i] for i in range(len(gshape))]
    return gshape, rows, np.ascontiguousarray(Q0), dX


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------
# algorithmic bytes
# ---------------------------------------------------------------------------
def algorithmic_model(cfg, shape, V, second_order, useB, useS):
    """Per-launch algorithmic bytes of each kernel (inputs read once + outputs written
    once) for a slab of `shape` interior cells, and B_alg per cell-update."""
    nd, N = len(shape), cfg['order']
    Nd, NP = N**nd, N**nd
    TRW = 1 + (nd if second_order else 0)
    WSW = 2 if second_order else 1
    FLXW = 2 if useB else 1
    D = 8
    cells = int(np.prod(shape))
    cw = int(np.prod([n + 2 for n in shape]))
    cb = int(np.prod([n + 2 * N for n in shape]))
    faces = [int(np.prod([n + (1 if i == d else 0) for i, n in enumerate(shape)]))
             for d in range(nd)]
    favg = sum(faces) / nd
    cen = V if (useB or useS) else 0
    rusanov = cfg['flux'] == 'rusanov'
    kb = {
        'k_boundaries': (cells + cb) * V * D,
        'k_weno2d': (cb + cw * Nd) * V * D,
        'k_cfl': cw * Nd * V * D,
        'k_dg': cw * (Nd * V + 2 * nd * NP * TRW * V + cen) * D,
        'k_wavespeeds': cw * 2 * nd * NP * (TRW * V + WSW) * D,
        'k_faces': favg * (2 * NP * (TRW * V + (WSW if rusanov else 0)) + FLXW * V) * D,
        'k_faces_fused': favg * (2 * NP * TRW * V + FLXW * V) * D,
        'k_update': (cells * V * 2 + sum(faces) * FLXW * V + cells * cen) * D,
    }
    kb['k_dg_n'] = kb['k_dg_stiff'] = kb['k_dg']
    kb['k_faces_side'] = kb['k_faces_fused']
    # k_weno_sweep: ndim launches of different sizes; the average launch
    sh = [n + 2 * N for n in shape]
    tot = 0
    for d in range(nd):
        n_in = int(np.prod(sh)) * N**d
        sh[d] -= 2 * (N - 1)
        tot += (n_in + int(np.prod(sh)) * N**(d + 1)) * V * D
    kb['k_weno_sweep'] = tot / nd
    kb['k_weno3d'] = (cb + cw * Nd) * V * D
    # SURVEY 8d three-product model: 8 V (3 + 2 Nd + 2 N Nd) bytes per cell-update
    b_alg = 8 * V * (3 + 2 * Nd + 2 * N * Nd)
    return kb, b_alg


def ncu_record(config, kernel):
    """Per-launch numbers of `kernel` from the committed `ncu --set full` capture of this
    configuration (profiles/r2_ncu_kernels.json, written by tools/ncu_summary.py): DRAM
    bytes, executed FP64 flops (2 DFMA + DMUL + DADD thread instructions), FP64 pipe %."""
    p = os.path.join(ROOT, 'profiles', 'r2_ncu_kernels.json')
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        d = json.load(f)
    rec = d.get(config, {}).get('kernels', {}).get(kernel)
    return rec, ('profiles/r2_ncu_kernels.json[%s] (%s)' % (config, d.get(config, {}).get('source')))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region.  The sampler is
    started early (nvidia-smi takes a second to come up) and runs through the whole
    bench; stop(t0, t1) keeps the samples whose time stamp lies in the wall-clock
    window [t0, t1] of the load (warm-up, timed steps, per-kernel pass)."""
    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '20'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self, t0=None, t1=None):
        import datetime
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(', ') for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        parsed = []
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), '%Y/%m/%d %H:%M:%S.%f').timestamp()
                rec = (ts, float(r[1]), float(r[2]), float(r[3]),
                       [nme for k, nme in enumerate(names)
                        if len(r) > 4 + k and r[4 + k].strip().lower() == 'active'])
            except (ValueError, IndexError):
                continue
            parsed.append(rec)
        inside = [r for r in parsed if t0 is None or (t0 - 0.02 <= r[0] <= t1 + 0.02)]
        if not inside and parsed and t0 is not None:
            # none landed inside the window: the sample nearest to it
            inside = [min(parsed, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))]
            out['note'] = 'no sample inside the load window; nearest sample reported'
        if inside:
            out['sm_mhz'] = float(np.median([r[1] for r in inside]))
            out['sm_max_mhz'] = inside[-1][2]
            out['power_w_max'] = max(r[3] for r in inside)
            out['samples'] = len(inside)
            out['window_s'] = None if t0 is None else round(t1 - t0, 3)
            out['reasons'] = sorted({x for r in inside for x in r[4]})
        return out


# ---------------------------------------------------------------------------
# reference CPU path (oracle/_ref = the unmodified reference, C callbacks)
# ---------------------------------------------------------------------------
def run_reference_sample(name, nsteps_target, threads):
    """Runs the reference's pde_solver on the bounded sample of configuration `name` for
    about nsteps_target steps; returns (cells*steps/s, steps, seconds, sample text)."""
    from oracle import reference as R
    cfg = CONFIGS[name]
    ref = cfg['ref']
    ndim = len(cfg['shape'])
    if 'sized' in ref:
        b = cases.sized_bases()[ref['sized']]
        Q0, L = b['Q0'], b['L']
        ks = sorted(b['tf'])
        k = min(ks, key=lambda x: abs(x - nsteps_target))
        tf = b['tf'][k]
    else:
        Q0, L = cfg['ic'](ref['shape']), cfg['L']
        tf = ref['tf'](nsteps_target)
    lib = 'libpypde_ref3d.so' if ndim == 3 else 'libpypde_ref.so'
    F, B, S = R.system_callbacks(cfg['system'], ndim)
    from pypde_b200.systems import SYSTEMS
    second = SYSTEMS[cfg['system']][4]
    # the reference prints "t = ..." per step on stdout (iterator.cpp:134): capture fd 1
    # to count the steps it took
    sys.stdout.flush()
    saved = os.dup(1)
    tmp = tempfile.TemporaryFile('w+b')
    os.dup2(tmp.fileno(), 1)
    t0 = time.perf_counter()
    try:
        R.pde_solver(Q0, tf, L, F=F, B=B, S=S, boundaryTypes=cfg['bts'], order=cfg['order'],
                     ndt=1, flux=cfg['flux'], stiff=cfg['stiff'], nThreads=threads,
                     secondOrder=second, lib=lib)
    finally:
        secs = time.perf_counter() - t0
        os.dup2(saved, 1)
        os.close(saved)
    tmp.seek(0)
    steps = sum(1 for l in tmp.read().decode().splitlines() if l.startswith('t = '))
    tmp.close()
    cells = int(np.prod(Q0.shape[:-1]))
    what = ('%s, %d steps in %.1f s, unmodified reference pde_solver (oracle/_ref%s) with C '
            'callbacks, nThreads=%d of %d host cores' %
            (ref['what'], steps, secs, ' + the one-line 3-D indexing fix' if ndim == 3 else '',
             threads, os.cpu_count() or 1))
    return cells * steps / secs, steps, secs, what


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import reference as R
    if not R.available('libpypde_ref.so'):
        print(json.dumps({'impl': 'reference', 'unavailable':
                          'oracle/_ref/libpypde_ref.so missing (run make -C oracle where '
                          '/root/reference exists)'}))
        return
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    threads = max(1, cores - 1)
    for _ in range(max(0, min(args.warmup, 1))):
        run_reference_sample(args.config, 1, threads)
    value, steps, secs, sample = run_reference_sample(args.config, args.steps, threads)
    line = {
        'impl': 'reference', 'metric': 'cell-updates/s', 'value': value, 'unit': 'cell-updates/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': secs / max(steps, 1) * 1e3, 'higher_is_better': True,
        'scaling': cfg['scaling'], 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': cfg['workload'], 'name': args.config, 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'cell-updates/s', 'cores': threads,
                         'kind': 'reference', 'sample': sample},
        'e2e': {'value': value, 'unit': 'cell-updates/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# multi-GPU correctness, outside the timed region: slabs reproduce the undivided run
# ---------------------------------------------------------------------------
SLAB_CHECKS = {
    # name: (system, shape, order, bts, tf, stiff) — every grid has >= 8 x order rows
    'euler2d': ('euler', (64, 48), 3, ['transitive', 'transitive'], 0.02, False),
    'euler2d_periodic': ('euler', (64, 40), 3, ['periodic', 'transitive'], 0.02, False),
    'ns3d_second_order': ('navier_stokes', (24, 6, 5), 3, ['periodic'] * 3, 0.03, False),
    'advect_nc_BS': ('advect_nc', (32, 12), 2, ['periodic', 'periodic'], 0.05, False),
    'gpr_stiff': ('gpr', (32, 8), 2, ['transitive', 'transitive'], 0.003, True),
}


def slab_check_ic(system, shape):
    return {'euler': cases.euler_smooth, 'navier_stokes': cases.taylor_green,
            'advect_nc': cases.advect_nc_smooth, 'gpr': cases.gpr_disc}[system](shape)


def slab_checks_undivided(names):
    """Rank 0, before the communicator exists: the undivided single-GPU runs."""
    import pypde_b200
    from pypde_b200.systems import cuda_sources
    out = {}
    for name in names:
        system, shape, N, bts, tf, stiff = SLAB_CHECKS[name]
        F, B, S, V = cuda_sources(system, len(shape))
        L = [2 * np.pi] * 3 if system == 'navier_stokes' else [1.] * len(shape)
        Q0 = slab_check_ic(system, shape)
        out[name] = pypde_b200.pde_solver(Q0.copy(), tf, L, F=F, B=B, S=S, boundaryTypes=bts,
                                          order=N, ndt=1, stiff=stiff)[0]
    return out


def slab_checks_divided(names, undivided, rank, world, dist):
    """All ranks, communicator up: each advances its slab through the pde_solver C ABI (NCCL
    halo exchange + dt all-reduce inside libpypde.so); rank 0 stitches and compares."""
    import pypde_b200
    from pypde_b200 import slabs
    from pypde_b200.systems import cuda_sources
    res = {}
    for name in names:
        system, shape, N, bts, tf, stiff = SLAB_CHECKS[name]
        F, B, S, V = cuda_sources(system, len(shape))
        Lg = [2 * np.pi] * 3 if system == 'navier_stokes' else [1.] * len(shape)
        Q0 = slab_check_ic(system, shape)
        mine = np.ascontiguousarray(slabs.split(Q0, world)[rank])
        # the slab keeps the global cell size: L_local = rows_local * dx
        Lloc = [Lg[0] * mine.shape[0] / shape[0]] + Lg[1:]
        out = pypde_b200.pde_solver(mine, tf, Lloc, F=F, B=B, S=S, boundaryTypes=bts, order=N,
                                    ndt=1, stiff=stiff)[0]
        gathered = [None] * world
        dist.all_gather_object(gathered, out)
        if rank == 0:
            st = slabs.stitch(gathered)
            res[name] = {'bit_identical': bool(np.array_equal(st, undivided[name])),
                         'max_abs_diff': float(np.abs(st - undivided[name]).max()),
                         'moved': float(np.abs(st - Q0).max()), 'grid': list(shape),
                         'order': N}
    return res


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=6)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS))
    ap.add_argument('--size', type=int, default=0,
                    help='override the cells per axis (per GPU where the config is weak-scaled)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-smooth', action='store_true',
                    help='c2: skip the smooth-periodic variant measured beside the explosion')
    ap.add_argument('--slab-checks', default='all',
                    help="multi-rank runs: comma-separated SLAB_CHECKS names, 'all' or 'none'")
    ap.add_argument('--user-functions', default='cuda', choices=['cuda', 'numba', 'traced'],
                    help='c2 / c2smooth only: the Euler flux as the CUDA text of '
                         'pypde_b200/systems (default), as a device-style Python function '
                         'lowered by numba-CUDA to LTO-IR, or as a reference-style Python '
                         'function lowered by tracing')
    ap.add_argument('--analytic-wavespeed', action='store_true',
                    help='OPT-IN experiment, never the default: analytic |v| + c through '
                         'pypde_b200_set_wavespeed instead of the reference-defined finite-'
                         'difference Jacobian eigen-solves (results differ at the 1e-8 level)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != 'reference' else args.warmup

    if args.impl == 'reference':
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from pypde_b200.handle import Solver, comm_init_from_torch
    from pypde_b200.systems import SYSTEMS, cuda_sources
    from pypde_b200.utils import create_solver, c_ptr, check_error

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — this framework has no CPU path')
    torch.cuda.set_device(local)
    torch.zeros(1, device='cuda')           # primary context
    os.environ['PYPDE_B200_QUIET'] = '1'
    cfg = CONFIGS[args.config]
    replica = cfg['mode'] == 'replica'
    slab_res = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        names = ([] if args.slab_checks == 'none' else
                 list(SLAB_CHECKS) if args.slab_checks == 'all' else args.slab_checks.split(','))
        names = [n for n in names if SLAB_CHECKS[n][1][0] >= world * SLAB_CHECKS[n][2]]
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):   # (pde_solver prints like the reference)
            undivided = slab_checks_undivided(names) if rank == 0 else None
        dist.barrier()
        if not replica or names:
            comm_init_from_torch()
        if names:
            with contextlib.redirect_stdout(sys.stderr):
                slab_res = slab_checks_divided(names, undivided, rank, world, dist)
        if replica:
            from pypde_b200.handle import _lib
            _lib().pypde_b200_comm_finalize()

    if args.analytic_wavespeed:
        from pypde_b200.systems import euler_wavespeed
        from pypde_b200.utils import get_cdll
        _ws = euler_wavespeed(2)
        assert get_cdll().pypde_b200_set_wavespeed(_ws.pointer) == 0
    sampler = ClockSampler(local) if rank == 0 else None
    K, W = args.steps, args.warmup
    ndim, N_ORDER = len(cfg['shape']), cfg['order']
    F, B, S, V = cuda_sources(cfg['system'], ndim)
    second = SYSTEMS[cfg['system']][4]
    if args.user_functions != 'cuda':
        if cfg['system'] != 'euler' or ndim != 2:
            raise SystemExit('bench.py: --user-functions applies to c2 / c2smooth')
        from pypde_b200.systems import python_functions
        F = python_functions.euler2d(args.user_functions)
    gshape, rows, Q0, dX = slab_problem(cfg, rank, world, args.size)
    shape = Q0.shape[:-1]
    bts = cfg['bts']
    kw = dict(F=F, B=B, S=S, boundaryTypes=bts, cfl=0.9, order=N_ORDER, dX=dX, flux=cfg['flux'],
              stiff=cfg['stiff'])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: state lives in HBM (a torch tensor), stepped in place
    u_dev = torch.from_numpy(Q0).cuda()
    sol = Solver(Q0.shape, None, **kw)
    # (a stream of its own, not the legacy default stream: small grids replay the step as a
    #  CUDA graph, and the legacy stream cannot be captured)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sol.set_stream(stream.cuda_stream)
    sol.bind_tensor(u_dev)
    sol.begin(1e9)
    # clocks are sampled from the warm-up through the timed region to the per-kernel
    # pass (the GPU is under the same load throughout)
    barrier()
    t_load0 = time.time()
    for _ in range(W):
        sol.step_async()
    barrier()
    l0 = sol.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        sol.step_async()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sol.launches - l0
    t_end, dt_last, nan = sol.sync()
    if nan or not np.isfinite(t_end):
        raise SystemExit('bench.py: NaNs in the solution')
    msr = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(msr, op=dist.ReduceOp.MAX)
    ms = float(msr.item())
    cells_rank = int(np.prod(shape))
    cells_total = cells_rank * world if cfg['mode'] != 'split' else int(np.prod(gshape))
    value = cells_total * K / (ms * 1e-3)

    # ---- per-kernel device times (CUDA events on the launching stream), separate pass
    sol.set_profiling(True)
    P = 5 if ms / K < 50. else 2
    for _ in range(P):
        sol.step_async()
    kt = sol.kernel_times()
    sol.set_profiling(False)
    t_load1 = time.time()
    clocks = sampler.stop(t_load0, t_load1) if sampler else None
    fp64_peak = sol.fp64_peak_tflops()
    hbm_peak, peak_src = measured_peaks()
    kb, b_alg = algorithmic_model(cfg, shape, V, second, B is not None, S is not None)
    step_ms_prof = sum(v_[0] for v_ in kt.values()) / P
    dom = max(kt, key=lambda k_: kt[k_][0])
    dom_ms = kt[dom][0] / kt[dom][1]            # average launch duration
    dom_gbs = kb.get(dom, 0.) / (dom_ms * 1e-3) / 1e9
    # DRAM bytes and executed FP64 flops per launch of that kernel from the committed
    # ncu --set full capture of this configuration (same grid only)
    rec, rec_src = ncu_record(args.config, dom)
    if rec is not None and (args.size or world > 1 and cfg['mode'] == 'split'):
        rec = None
    hbm = {'achieved': dom_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': dom_gbs / hbm_peak,
           'algorithmic_bytes_per_launch': kb.get(dom), 'peak_source': peak_src}
    fp64 = None
    if rec is not None and rec.get('fp64_flops') and fp64_peak:
        tfl = rec['fp64_flops'] / (dom_ms * 1e-3) / 1e12
        fp64 = {'achieved': tfl, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': tfl / fp64_peak,
                'flops_per_launch': rec['fp64_flops'],
                'flops_source': '2 x DFMA + DMUL + DADD thread instructions executed '
                                '(smsp__sass_thread_inst_executed_op_d{fma,mul,add}_pred_on), ' +
                                rec_src,
                'pipe_fp64_pct': rec.get('pipe_fp64_pct'),
                'peak_source': 'DFMA micro-kernel measured in this run (k_fp64_peak)'}
    # the bound that binds: the FP64 pipe's busy fraction (ncu) against the algorithmic
    # share of the HBM bandwidth
    pipe = (rec or {}).get('pipe_fp64_pct')
    fp64_bound = pipe is not None and pipe / 100. > hbm['frac']
    if fp64_bound and fp64:
        roofline = {'kernel': dom, 'bound': 'fp64', 'achieved': fp64['achieved'],
                    'peak': fp64['peak'], 'unit': 'TFLOP/s', 'frac': fp64['frac']}
    else:
        roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': hbm['achieved'],
                    'peak': hbm['peak'], 'unit': 'GB/s', 'frac': hbm['frac']}
    roofline.update({
        'traffic': (rec or {}).get('dram_bytes'), 'traffic_source': rec_src if rec else None,
        'avg_launch_ms': dom_ms, 'launches_per_step': kt[dom][1] / P,
        'share_of_step': kt[dom][0] / P / step_ms_prof,
        'hbm': hbm, 'fp64': fp64,
        'kernels_ms_per_step': {k_: kt[k_][0] / P for k_ in kt},
        'step': {'b_alg_bytes_per_cell_update': b_alg,
                 'hbm_frac': value / world * b_alg / 1e9 / hbm_peak}})
    sol.close()
    del u_dev

    # ---- the same workload on SURVEY 8d's smooth periodic data (no plateaus: two Picard
    # iterations in every cell, no root search that ends in one step), device-resident, in the
    # same run: the default line carries both numbers
    smooth = None
    if args.config == 'c2' and world == 1 and not args.no_smooth:
        scfg = CONFIGS['c2smooth']
        _, _, Qs, dXs = slab_problem(scfg, 0, 1, args.size)
        us = torch.from_numpy(Qs).cuda()
        ss = Solver(Qs.shape, None, F=F, boundaryTypes=scfg['bts'], cfl=0.9, order=N_ORDER,
                    dX=dXs, flux=scfg['flux'], stiff=False)
        ss.set_stream(stream.cuda_stream)
        ss.bind_tensor(us)
        ss.begin(1e9)
        for _ in range(W):
            ss.step_async()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(K):
            ss.step_async()
        s1.record(stream)
        torch.cuda.synchronize()
        sms = s0.elapsed_time(s1)
        _, _, snan = ss.sync()
        ss.close()
        del us
        if not snan:
            smooth = {'value': cells_rank * K / (sms * 1e-3), 'unit': 'cell-updates/s',
                      'ms_per_step': sms / K, 'workload': scfg['workload']}

    # ---- end-to-end arm: the reference-facing C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        # final time after exactly K steps from the IC (dt is global, so identical on all ranks)
        probe = Solver(Q0.shape, None, **kw)
        probe.set_state(Q0)
        probe.begin(1e9)
        for _ in range(K):
            probe.step_async()
        tf_k, _, _ = probe.sync()
        probe.close()
        solver = create_solver()
        nX = np.array(shape, dtype='int32')
        dXa = np.array(dX)
        bt = np.array([{'transitive': 0, 'periodic': 1}[b] for b in bts], dtype='int32')
        flux_id = {'rusanov': 0, 'roe': 1, 'osher': 2}[cfg['flux']]
        pinned_u = torch.from_numpy(Q0.copy()).pin_memory()
        pinned_ret = torch.zeros(Q0.size, dtype=torch.float64).pin_memory()
        ur, ret = pinned_u.numpy().ravel(), pinned_ret.numpy()

        def call():
            solver(F.ctypes, B.ctypes if B else None, S.ctypes if S else None, True,
                   B is not None, S is not None, c_ptr(ur), tf_k, c_ptr(nX), ndim, c_ptr(dXa),
                   0.9, c_ptr(bt), cfg['stiff'], flux_id, N_ORDER, V, 1, second, c_ptr(ret), 1)
            check_error('pde_solver')

        call()                                   # warm: JIT cache, allocator
        ur[:] = Q0.ravel()
        barrier()
        t0 = time.perf_counter()
        call()
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        st = torch.tensor([secs], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(st, op=dist.ReduceOp.MAX)
        secs = float(st.item())
        e2e = {'value': cells_total * K / secs, 'unit': 'cell-updates/s',
               'h2d_bytes_per_step': Q0.nbytes / K,
               'd2h_bytes_per_step': (Q0.nbytes + 56 * K) / K,
               'seconds': secs, 'steps': K,
               'what': 'one pde_solver() C-ABI call from pinned host Q0 to host ret for exactly K '
                       'steps (second call: the library keeps the solver — kernel module and '
                       'work arrays — of the previous call with the same configuration): H2D of '
                       'Q0, K steps with a per-step sync + D2H of (t, dt), one D2H of the final '
                       'state into Q0 and a host copy of it into ret'}

    # ---- CPU baseline beside it (rank 0, bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import reference as R
            if R.available('libpypde_ref.so'):
                cores = os.cpu_count() or 1
                threads = max(1, cores - 1)
                val, steps, secs, sample = run_reference_sample(args.config, 10, threads)
                cpu = {'value': val, 'unit': 'cell-updates/s', 'cores': threads,
                       'kind': 'reference', 'sample': sample}
        except Exception as ex:  # the baseline must never take the bench line down
            cpu = {'value': None, 'unit': 'cell-updates/s', 'cores': 0, 'kind': 'reference',
                   'sample': 'failed: %r' % (ex, )}

    if rank == 0:
        Nd = N_ORDER**ndim
        TRW = 1 + (ndim if second else 0)
        cw = int(np.prod([n + 2 for n in shape]))
        line = {
            'metric': 'cell-updates/s', 'value': value, 'unit': 'cell-updates/s',
            'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms / K,
            'higher_is_better': True, 'scaling': cfg['scaling'], 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': cfg['workload'], 'name': args.config,
                       'cells_per_gpu': cells_rank, 'global_grid': list(gshape),
                       'order': N_ORDER, 'flux': cfg['flux'], 'stiff': cfg['stiff'],
                       'user_functions': args.user_functions,
                       'l2': 'inputs larger than L2: w + traces = %.2f GB per step vs 126 MB L2'
                             % (cw * (Nd * V + 2 * ndim * Nd * TRW * V) * 8 / 1e9)
                             if cw * (Nd * V + 2 * ndim * Nd * TRW * V) * 8 > 126e6 else
                             'the whole working set (%.1f MB) fits L2: the case is what '
                             'BASELINE names, latency bound by construction'
                             % (cw * (Nd * V + 2 * ndim * Nd * TRW * V) * 8 / 1e6),
                       'parallelism': ('replicas%d' if replica else 'slab%d') % world,
                       **({'wavespeed': 'analytic |v|+c via pypde_b200_set_wavespeed — OPT-IN '
                                        'experiment, NOT the reference-defined path (its wave '
                                        'speeds are spectral radii of finite-difference '
                                        'Jacobians); not comparable with the default line'}
                          if args.analytic_wavespeed else {})},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline,
            'cpu_baseline': cpu,
        }
        if smooth is not None:
            line['smooth_periodic_variant'] = smooth
        if world > 1:
            line['slab_bit_identical'] = (None if not slab_res else
                                          all(v['bit_identical'] for v in slab_res.values()))
            line['slab_checks'] = slab_res
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
