"""Prints the GPU-vs-reference parity of every golden case (relative L-inf) next to the
reference's own +-1 ulp self-noise (maximum over four seeds) and the tolerance the tests
apply: max(stated, 4 x self-noise); copy the output into profiles/."""
import sys

import numpy as np

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
sys.path.insert(0, 'tests/golden')
import cases
from conftest import parity_tolerance, rel_linf
from test_gpu_parity import run_gpu

print('%-32s %-12s %-12s %-12s %-12s %s' % ('case', 'grid', 'GPU vs ref', 'self-noise', 'tolerance',
                                             'err/noise'))
for fname, table in (('solver', cases.solver_cases()), ('solver_sized', cases.sized_cases())):
    g = np.load('tests/golden/%s.npz' % fname)
    for name, c in table.items():
        out, _ = run_gpu(c)
        err, noise = rel_linf(out[0], g[name]), float(g[name + '__noise'])
        tol = parity_tolerance(g, name, 1e-8 if c.get('stiff') else 1e-10)
        print('%-32s %-12s %-12.3e %-12.2e %-12.2e %-6.2f %s' % (
            name, 'x'.join(map(str, c['Q0'].shape[:-1])), err, noise, tol, err / noise,
            '' if err < tol else 'ABOVE TOLERANCE'), flush=True)
