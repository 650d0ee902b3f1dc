"""tools/hot_lines.py for a kernel of a bench.py configuration: rebuilds the configuration's
cubin (the JIT is deterministic and cached) and attributes the executed instructions and
stall samples of an ncu report to source lines.

    python tools/hot_config.py <report.ncu-rep> <config> <kernel> [top]
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import bench  # noqa: E402
from pypde_b200.systems import SYSTEMS  # noqa: E402
from sass_stats import build_cubin  # noqa: E402

rep, name, kernel = sys.argv[1:4]
top = sys.argv[4] if len(sys.argv) > 4 else '40'
cfg = bench.CONFIGS[name]
ndim = len(cfg['shape'])
vfun, hasF, hasB, hasS, second = SYSTEMS[cfg['system']]
flux = {'rusanov': 0, 'roe': 1, 'osher': 2}[cfg['flux']]
cubin = build_cubin(cfg['system'], ndim, cfg['order'], flux, int(cfg['stiff']))
with tempfile.NamedTemporaryFile(suffix='.cubin', delete=False) as f:
    f.write(cubin)
env = dict(os.environ, HOT_LINES_CFG='%d,%d,%d,%d,%d,%d,%d' % (
    ndim, cfg['order'], vfun(ndim), int(hasF), int(hasB), int(hasS), int(second)))
subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'hot_lines.py'), rep, f.name, kernel,
                top], env=env)
os.unlink(f.name)
