"""Attributes the per-instruction execution counts of an ncu report (source
page, SASS) to lines of the specialised CUDA source, using the line table of
the identical cubin built locally (`nvdisasm -g`).

    python tools/hot_lines.py <report.ncu-rep> <cubin> <kernel> [top]

The cubin must be the one the report was taken from (tools/sass_stats.py: build_cubin
rebuilds it here — the JIT is deterministic).  HOT_LINES_CFG=ndim,N,V,useF,useB,useS,secondOrder
selects the specialised source whose lines are printed (default 2,3,4,1,0,0,0 = BASELINE
configs[1]).
"""
import collections
import csv
import ctypes
import io
import re
import subprocess
import sys

sys.path.insert(0, '.')


def ncu_instr(rep, kernel):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, res, want_hdr = None, None, [], False
    for r in rows:
        if len(r) >= 2 and r[0] == 'Kernel Name':
            if res:          # only the first captured launch of this kernel
                break
            cur, want_hdr = r[1], True
            continue
        if cur != kernel:
            continue
        if want_hdr:
            hdr, want_hdr = r, False
            continue
        res.append(r)
    return hdr, res


def line_table(cubin, kernel):
    out = subprocess.run(['nvdisasm', '-g', cubin], capture_output=True, text=True).stdout
    res, cur, active = [], None, False
    for l in out.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
        if m:
            active = m.group(1) == kernel
            continue
        if re.match(r'\s*\.section', l):
            active = False
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            # device functions of other translation units (the user's F/B/S) keep their
            # own file name; report them as "<file>:<line>"
            f = m.group(1)
            cur = int(m.group(2)) if f.endswith('pypde_b200_kernels.cu') else '%s:%s' % (
                f.split('/')[-1], m.group(2))
            continue
        m = re.match(r'\s*(\$\S+):', l)
        if m:   # compiler-provided subroutine (div / sqrt slow paths): no line info
            cur = m.group(1).split('$')[-1]
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+[A-Z@]', l):
            res.append((cur, l.strip()))
    return res


def source_lines(ndim, N, V, useF=1, useB=0, useS=0, so=0):
    from pypde_b200.utils import get_cdll
    lib = get_cdll()
    lib.pypde_b200_emit_source.argtypes = [ctypes.c_int] * 9 + [ctypes.c_char_p, ctypes.c_size_t,
                                                                ctypes.POINTER(ctypes.c_size_t)]
    buf = ctypes.create_string_buffer(1 << 21)
    n = ctypes.c_size_t()
    lib.pypde_b200_emit_source(ndim, N, V, 0, 0, useF, useB, useS, so, buf, 1 << 21, ctypes.byref(n))
    src = buf.value.decode().splitlines()
    k = 0
    while src[k].startswith('#define PDE_'):
        k += 1
    return src[k:]


if __name__ == '__main__':
    rep, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    hdr, rows = ncu_instr(rep, kernel)
    lt = line_table(cubin, kernel)
    import os
    _c = [int(x) for x in os.environ.get("HOT_LINES_CFG", "2,3,4,1,0,0,0").split(",")]
    body = source_lines(*_c)
    ie = hdr.index('Instructions Executed')
    isamp = hdr.index('# Samples')
    print('%s: %d SASS instructions in report, %d in local cubin' % (kernel, len(rows), len(lt)))
    if len(rows) != len(lt):
        sys.exit('line table does not match the report (different build?)')
    per, samp, tot, stot = collections.Counter(), collections.Counter(), 0, 0
    ops = collections.Counter()
    for (ln, txt), r in zip(lt, rows):
        c, s = int(r[ie]), int(r[isamp])
        per[ln] += c
        samp[ln] += s
        tot += c
        stot += s
        m = re.match(r'/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)', txt)
        if m:
            ops[m.group(2)] += c
    print('total warp instructions executed: %d; stall samples: %d' % (tot, stot))
    print('--- by opcode')
    for k, v in ops.most_common(14):
        print('  %5.1f%%  %s' % (100. * v / tot, k))
    print('--- by source line (share of executed instructions | share of stall samples)')
    for ln, c in per.most_common(top):
        text = (body[ln - 1].strip()[:96] if isinstance(ln, int) and ln - 1 < len(body)
                else ('(user function)' if ':' in str(ln) else '(compiler subroutine)'))
        print('  %5.1f%% %5.1f%%  line %4s: %s' % (100. * c / tot, 100. * samp[ln] / max(stot, 1), ln,
                                                 text))
