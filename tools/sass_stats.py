"""Static SASS statistics of one specialised cubin (no GPU needed): registers,
instruction count and the expensive opcodes (MUFU = division / sqrt seeds, BRA,
local-memory traffic) per kernel.  Used to compare kernel variants on the CPU
box before spending GPU time.

    [PYPDE_B200_EXTRA_DEFINES=..] python tools/sass_stats.py [system ndim N flux stiff] [kernel ...]
"""
import collections
import ctypes
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build_cubin(system='euler', ndim=2, N=3, flux=0, stiff=0):
    from pypde_b200.systems import cuda_sources
    from pypde_b200.utils import get_cdll, last_error
    lib = get_cdll()
    F, B, S, V = cuda_sources(system, ndim)
    so = bool(getattr(F, 'second_order', False))
    n = ctypes.c_size_t()
    args = (F.pointer if F else None, B.pointer if B else None, S.pointer if S else None, ndim, N,
            V, flux, stiff, int(so))
    if lib.pypde_b200_compile(*args, ctypes.byref(n), None, ctypes.c_size_t(0)) != 0:
        raise RuntimeError(last_error())
    buf = ctypes.create_string_buffer(n.value)
    if lib.pypde_b200_compile(*args, ctypes.byref(n), buf, ctypes.c_size_t(n.value)) != 0:
        raise RuntimeError(last_error())
    return buf.raw[:n.value]


def stats(cubin_bytes, kernels=None):
    with tempfile.NamedTemporaryFile(suffix='.cubin', delete=False) as f:
        f.write(cubin_bytes)
        path = f.name
    try:
        sass = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
        res = subprocess.run(['cuobjdump', '-res-usage', path], capture_output=True, text=True).stdout
    finally:
        os.unlink(path)
    regs = {}
    cur = None
    for l in res.splitlines():
        m = re.search(r'Function (\S+):', l)
        if m:
            cur = m.group(1)
        m = re.search(r'REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)', l)
        if m and cur:
            regs[cur] = tuple(int(x) for x in m.groups())
    out = {}
    cur = None
    for l in sass.splitlines():
        m = re.search(r'Function : (\S+)', l)
        if m:
            cur = m.group(1)
            out[cur] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
        if m and cur:
            op = m.group(2)
            out[cur]['total'] += 1
            out[cur][op.split('.')[0]] += 1
            if op.startswith('MUFU'):
                out[cur][op] += 1
    for k, c in out.items():
        if kernels and k not in kernels:
            continue
        r = regs.get(k, (0, 0, 0))
        mufu = ' '.join('%s=%d' % (a, b) for a, b in sorted(c.items()) if a.startswith('MUFU.'))
        print('%-16s regs=%3d local=%4d  instr=%6d  DFMA=%d DMUL=%d DADD=%d BRA=%d CALL=%d LDL=%d STL=%d  %s' %
              (k, r[0], r[2], c['total'], c['DFMA'], c['DMUL'], c['DADD'], c['BRA'], c['CALL'],
               c['LDL'], c['STL'], mufu))


if __name__ == '__main__':
    a = sys.argv[1:]
    cfg = ('euler', 2, 3, 0, 0)
    if len(a) >= 5 and a[1].isdigit():
        cfg = (a[0], int(a[1]), int(a[2]), int(a[3]), int(a[4]))
        a = a[5:]
    stats(build_cubin(*cfg), a or None)
