"""Prints the GPU-vs-reference parity of every golden case (relative L-inf)
next to the oracle-vs-reference figure; copy the output into profiles/."""
import sys

import numpy as np

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
sys.path.insert(0, 'tests/golden')
import cases
from conftest import rel_linf
from test_gpu_parity import run_gpu

g = np.load('tests/golden/solver.npz')
print('%-26s %-12s %-26s %s' % ('case', 'grid', 'rel Linf GPU vs reference', 'reference +-1ulp self-noise'))
for name, c in cases.solver_cases().items():
    out, _ = run_gpu(c)
    print('%-26s %-12s %-26.3e %.2e' % (name, 'x'.join(map(str, c['Q0'].shape[:-1])),
                                        rel_linf(out[0], g[name]), float(g[name + '__noise'])))
