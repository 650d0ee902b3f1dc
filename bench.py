#!/usr/bin/env python
"""bench.py — cell-updates/s of the ADER-WENO step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]           (our CUDA path)
    python bench.py --impl reference [--steps K] [--warmup W]     (reference CPU path)

Workload (config.workload): BASELINE.json configs[1] — 2-D Euler, cylindrical
explosion, 2048^2 cells per GPU, order 3, Rusanov flux, FP64, transmissive
boundaries.  With N > 1 ranks (one process per GPU under torchrun) the domain is
(N*2048) x 2048, slab-partitioned along axis 0 (weak scaling): every step does a
real NCCL halo exchange of N rows of u per side and a max-all-reduce for dt.

A step is one pass of the hot path (ghost cells, WENO, CFL/dt, DG predictor,
interface fluxes, update).  `value` is measured with the state resident in HBM;
`e2e` goes through the reference-facing C ABI (`pde_solver`) with host buffers.
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

N_ORDER = 3
V = 4
NDIM = 2


def explosion_slab(n):
    import cases
    return cases.euler_explosion((n, n))


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------
# algorithmic bytes / flops
# ---------------------------------------------------------------------------
def algorithmic_model(n):
    N, Nd, NP = N_ORDER, N_ORDER**NDIM, N_ORDER * N_ORDER**(NDIM - 1)
    cells = n * n
    cw = (n + 2) * (n + 2)
    D = 8
    # per-launch algorithmic bytes of each kernel = its inputs read once + outputs written once
    kb = {
        'k_boundaries': (cells + (n + 2 * N)**2) * V * D,
        'k_weno_sweep': None,  # two launches with different sizes, summed below
        'k_cfl': cw * Nd * V * D,
        'k_dg': cw * (Nd * V + 2 * NDIM * NP * V) * D,
        'k_dg_n': cw * (Nd * V + 2 * NDIM * NP * V) * D,             # w in, face traces out
        'k_faces_fused': n * (n + 1) * (2 * NP * V + V) * D,         # per direction: traces in, flux out
        'k_faces_side': n * (n + 1) * (2 * NP * V + V) * D,          # the same, two threads per face
        'k_wavespeeds': cw * 2 * NDIM * NP * (V + 1) * D,            # traces in, lambda out
        'k_faces': n * (n + 1) * (2 * NP * (V + 1) + V) * D,         # per direction
        'k_update': cells * V * D * 2 + NDIM * n * (n + 1) * V * D,
    }
    s0 = ((n + 2 * N)**2 + (n + 2) * (n + 2 * N) * N) * V * D
    s1 = ((n + 2) * (n + 2 * N) * N + cw * Nd) * V * D
    kb['k_weno_sweep'] = (s0 + s1) / 2.
    kb['k_weno2d'] = ((n + 2 * N)**2 + cw * Nd) * V * D                # ub in, w out
    # SURVEY 8d three-product model: 8 V (3 + 2 Nd + 2 N Nd) bytes per cell-update
    b_alg = 8 * V * (3 + 2 * Nd + 2 * N * Nd)
    f_alg = 4.5e4   # flop per cell-update at this config, ~70% in the face eigen-solves
    f_faces = 3.1e4          # per cell-update: the face eigen-solves + fluxes (k_wavespeeds)
    return kb, b_alg, f_alg, f_faces


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
def run_reference_sample(n, nsteps_target, threads):
    """Runs the reference's pde_solver on an n x n explosion for about
    nsteps_target steps; returns (cells*steps/s, steps, seconds)."""
    from oracle import reference as R
    import cases
    Q0 = cases.euler_explosion((n, n))
    F, B, S = R.system_callbacks('euler', 2)

    def solve(tf):
        # the reference prints "t = ..." per step on stdout (iterator.cpp:134):
        # capture fd 1 to count the steps it took
        sys.stdout.flush()
        saved = os.dup(1)
        tmp = tempfile.TemporaryFile('w+b')
        os.dup2(tmp.fileno(), 1)
        t0 = time.perf_counter()
        try:
            R.pde_solver(Q0, tf, [1., 1.], F=F, order=N_ORDER, ndt=1, stiff=False,
                         nThreads=threads)
        finally:
            dt = time.perf_counter() - t0
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        lines = [l for l in tmp.read().decode().splitlines() if l.startswith('t = ')]
        tmp.close()
        return len(lines), dt, [float(l[4:]) for l in lines]

    # explosion IC at rest: lambda_max = c = sqrt(1.4), dx = 1/n; the first 6 steps
    # use 0.2 dt (stepper.cpp:69-70)
    dt0 = 0.2 * 0.9 / (2 * np.sqrt(1.4) * n)
    k = nsteps_target
    tf = (min(k, 6) + 5 * max(0, k - 6)) * dt0 * 0.999
    steps, secs, _ = solve(tf)
    return n * n * steps / secs, steps, secs


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
    cores = os.cpu_count() or 1
    threads = max(1, cores - 1)
    n = args.ref_size
    for _ in range(max(0, min(args.warmup, 1))):
        run_reference_sample(n, 1, threads)
    value, steps, secs = run_reference_sample(n, args.steps, threads)
    sample = ('%dx%d cells of the same workload (2-D Euler explosion, order 3, Rusanov), %d steps, '
              'reference pde_solver with C callbacks, nThreads=%d of %d host cores' %
              (n, n, steps, threads, cores))
    line = {
        'impl': 'reference', 'metric': 'cell-updates/s', 'value': value, 'unit': 'cell-updates/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': secs / max(steps, 1) * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'cell-updates/s', 'cores': threads,
                         'kind': 'reference', 'sample': sample},
        'e2e': {'value': value, 'unit': 'cell-updates/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


WORKLOAD = ('2-D Euler cylindrical explosion, 2048^2 cells per GPU, order 3, Rusanov, FP64, '
            'transmissive (BASELINE configs[1]); slab-partitioned along axis 0 for N>1')


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=6)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--size', type=int, default=2048, help='cells per axis per GPU')
    ap.add_argument('--ref-size', type=int, default=128, help='grid of the CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
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
    from pypde_b200.systems import cuda_sources
    from pypde_b200.utils import create_solver, c_ptr, check_error

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device — this framework has no CPU path')
    torch.cuda.set_device(local)
    torch.zeros(1, device='cuda')           # primary context
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        comm_init_from_torch()

    os.environ['PYPDE_B200_QUIET'] = '1'
    if args.analytic_wavespeed:
        from pypde_b200.systems import euler_wavespeed
        from pypde_b200.utils import get_cdll
        _ws = euler_wavespeed(2)
        assert get_cdll().pypde_b200_set_wavespeed(_ws.pointer) == 0
    sampler = ClockSampler(local) if rank == 0 else None
    n = args.size
    K, W = args.steps, args.warmup
    F, B, S, v = cuda_sources('euler', 2)
    assert v == V
    Q0 = explosion_slab(n)
    dX = [1. / n, 1. / n]
    bts = ['transitive', 'transitive']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: state lives in HBM (a torch tensor), stepped in place
    u_dev = torch.from_numpy(Q0).cuda()
    sol = Solver(Q0.shape, None, F=F, boundaryTypes=bts, cfl=0.9, order=N_ORDER, dX=dX)
    stream = torch.cuda.current_stream()
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
    cells_total = n * n * world
    value = cells_total * K / (ms * 1e-3)

    # ---- per-kernel device times (CUDA events on the launching stream), separate pass
    sol.set_profiling(True)
    P = 5
    for _ in range(P):
        sol.step_async()
    kt = sol.kernel_times()
    sol.set_profiling(False)
    t_load1 = time.time()
    clocks = sampler.stop(t_load0, t_load1) if sampler else None
    fp64_peak = sol.fp64_peak_tflops()
    hbm_peak, peak_src = measured_peaks()
    kb, b_alg, f_alg, f_faces = algorithmic_model(n)
    step_ms_prof = sum(v_[0] for v_ in kt.values()) / P
    dom = max(kt, key=lambda k_: kt[k_][0])
    dom_ms = kt[dom][0] / kt[dom][1]            # average launch duration
    dom_gbs = kb.get(dom, 0.) / (dom_ms * 1e-3) / 1e9
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', 'r1_ncu_traffic.json')
    if n == 2048 and os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        if dom in tj.get('kernels', {}):
            traffic = tj['kernels'][dom]['dram_bytes']
            traffic_src = 'profiles/r1_ncu_traffic.json (ncu --set full, dram__bytes_read+write)'
    roofline = {
        'kernel': dom, 'bound': 'hbm', 'achieved': dom_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
        'frac': dom_gbs / hbm_peak, 'traffic': traffic, 'traffic_source': traffic_src,
        'algorithmic_bytes_per_launch': kb.get(dom), 'peak_source': peak_src,
        'avg_launch_ms': dom_ms, 'share_of_step': kt[dom][0] / P / step_ms_prof,
        'note': ('%s is FP64-pipe bound (finite-difference Jacobians + eigen-solves per face '
                 'node), not HBM bound; fp64 figures below' % dom),
        'fp64': {'achieved_tflops': (f_faces * n * n / (dom_ms * 1e-3) / 1e12)
                 if dom == 'k_wavespeeds' else
                 ((f_faces / NDIM) * n * n / (dom_ms * 1e-3) / 1e12
                  if dom in ('k_faces_fused', 'k_faces_side')
                  else None),
                 'peak_tflops': fp64_peak, 'peak_source': 'measured DFMA micro-kernel (k_fp64_peak)'},
        'kernels_ms_per_step': {k_: kt[k_][0] / P for k_ in kt},
        'step': {'b_alg_bytes_per_cell_update': b_alg,
                 'hbm_frac': value / world * b_alg / 1e9 / hbm_peak,
                 'f_alg_flop_per_cell_update': f_alg,
                 'fp64_frac': value / world * f_alg / 1e12 / fp64_peak if fp64_peak else None},
    }
    if roofline['fp64']['achieved_tflops'] and fp64_peak:
        roofline['fp64']['frac'] = roofline['fp64']['achieved_tflops'] / fp64_peak
    sol.close()
    del u_dev

    # ---- end-to-end arm: the reference-facing C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        # final time after exactly K steps from the IC (dt is global, so identical on all ranks)
        probe = Solver(Q0.shape, None, F=F, boundaryTypes=bts, cfl=0.9, order=N_ORDER, dX=dX)
        probe.set_state(Q0)
        probe.begin(1e9)
        for _ in range(K):
            probe.step_async()
        tf_k, _, _ = probe.sync()
        probe.close()
        solver = create_solver()
        nX = np.array([n, n], dtype='int32')
        dXa = np.array(dX)
        bt = np.array([0, 0], dtype='int32')
        pinned_u = torch.from_numpy(Q0.copy()).pin_memory()
        pinned_ret = torch.zeros(Q0.size, dtype=torch.float64).pin_memory()
        ur, ret = pinned_u.numpy().ravel(), pinned_ret.numpy()

        def call():
            solver(F.ctypes, None, None, True, False, False, c_ptr(ur), tf_k, c_ptr(nX), 2,
                   c_ptr(dXa), 0.9, c_ptr(bt), False, 0, N_ORDER, V, 1, False, c_ptr(ret), 1)
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
               'd2h_bytes_per_step': (2 * Q0.nbytes + 56 * K) / K,
               'seconds': secs, 'steps': K,
               'what': 'one pde_solver() C-ABI call from pinned host Q0 to host ret for exactly K '
                       'steps (second call: the library keeps the solver — kernel module and '
                       'work arrays — of the previous call with the same configuration): H2D of '
                       'Q0, K steps with a per-step sync + D2H of (t, dt), D2H of the final '
                       'state into ret and Q0'}

    # ---- CPU baseline beside it (rank 0, bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import reference as R
            if R.available('libpypde_ref.so'):
                cores = os.cpu_count() or 1
                threads = max(1, cores - 1)
                val, steps, secs = run_reference_sample(args.ref_size, 12, threads)
                cpu = {'value': val, 'unit': 'cell-updates/s', 'cores': threads,
                       'kind': 'reference',
                       'sample': '%dx%d cells of the same workload, %d steps in %.1f s, reference '
                                 'pde_solver (oracle/_ref) with C callbacks, nThreads=%d of %d host '
                                 'cores' % (args.ref_size, args.ref_size, steps, secs, threads,
                                            cores)}
        except Exception as ex:  # the baseline must never take the bench line down
            cpu = {'value': None, 'unit': 'cell-updates/s', 'cores': 0, 'kind': 'reference',
                   'sample': 'failed: %r' % (ex, )}

    if rank == 0:
        line = {
            'metric': 'cell-updates/s', 'value': value, 'unit': 'cell-updates/s',
            'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms / K,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'cells_per_gpu': n * n, 'order': N_ORDER,
                       'flux': 'rusanov',
                       'l2': 'inputs larger than L2: w + traces = %.1f GB per step vs 126 MB L2'
                             % ((n + 2)**2 * (9 * 4 + 144) * 8 / 1e9),
                       'parallelism': 'slab%d' % world,
                       **({'wavespeed': 'analytic |v|+c via pypde_b200_set_wavespeed — OPT-IN '
                                        'experiment, NOT the reference-defined path (its wave '
                                        'speeds are spectral radii of finite-difference '
                                        'Jacobians); not comparable with the default line'}
                          if args.analytic_wavespeed else {})},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline,
            'cpu_baseline': cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
