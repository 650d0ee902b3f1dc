"""Test tooling (build container only): prints, for the survey-sized parity cases of
tests/golden/cases.py, the times t_k the UNMODIFIED reference (oracle/_ref) reaches
after each of its first steps, and final times tf_K in the middle of step K (so that a
run to tf_K takes exactly K steps, the last one clipped: stepper.cpp:72-73) for
K = 1, 5, 10.  The values are pasted into cases.sized_cases()."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import cases  # noqa: E402
from oracle import reference as R  # noqa: E402


def step_times(c, tf, threads):
    ndim = c['Q0'].ndim - 1
    lib = 'libpypde_ref3d.so' if ndim == 3 else 'libpypde_ref.so'
    F, B, S = R.system_callbacks(c['system'], ndim)
    sys.stdout.flush()
    saved = os.dup(1)
    tmp = tempfile.TemporaryFile('w+b')
    os.dup2(tmp.fileno(), 1)
    t0 = time.perf_counter()
    try:
        R.pde_solver(c['Q0'], tf, c['L'], F=F, B=B, S=S, boundaryTypes=c['bts'],
                     order=c['order'], ndt=1, flux=c.get('flux', 'rusanov'),
                     stiff=c.get('stiff', False), nThreads=threads,
                     secondOrder=c.get('second_order', False), lib=lib)
    finally:
        secs = time.perf_counter() - t0
        os.dup2(saved, 1)
        os.close(saved)
    tmp.seek(0)
    ts = [float(l[4:]) for l in tmp.read().decode().splitlines() if l.startswith('t = ')]
    tmp.close()
    return ts, secs


def main():
    threads = max(1, (os.cpu_count() or 2) - 1)
    for name, c in cases.sized_bases().items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        # one clipped step to a tiny tf costs a full step but tells nothing; instead run
        # 2 steps' worth from a guess and grow until 10 steps are seen
        tf = c['tf_guess']
        while True:
            ts, secs = step_times(c, tf, threads)
            if len(ts) > 10:
                break
            tf *= 11.5 / max(len(ts) - 0.5, 0.5)
        ts = ts[:11]
        mids = {K: 0.5 * (ts[K - 2] if K > 1 else 0.) + 0.5 * ts[K - 1] for K in (1, 5, 10)}
        print('%-20s %.1f s for %d steps; t_k = %s' % (name, secs, len(ts), ts))
        print('    tf = {%s}' % ', '.join('%d: %.6g' % (K, mids[K]) for K in (1, 5, 10)))


if __name__ == '__main__':
    main()
