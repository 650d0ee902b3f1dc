"""The C-ABI library on a machine without a GPU: it loads, exports every
symbol include/pypde_b200.h declares, JIT-builds sm_100a cubins, and every
compute entry point fails loudly instead of falling back to the CPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import pypde_b200
from pypde_b200.cfuncs import CudaSource
from pypde_b200.systems import SYSTEMS, cuda_sources
from pypde_b200.utils import get_cdll, last_error, lib_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'pypde_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'\b(pypde_b200_\w+|pde_solver|weno_solver)\s*\(', text)
    return sorted(set(n for n in names if n != 'pypde_b200_devfn'))


def test_library_exports_every_declared_symbol():
    lib = get_cdll()
    names = header_functions()
    assert 'pde_solver' in names and 'weno_solver' in names and len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_reference_abi_symbols_are_unmangled_c():
    out = subprocess.check_output(['nm', '-D', '--defined-only', lib_path()]).decode()
    syms = {l.split()[-1] for l in out.splitlines() if l.strip()}
    assert {'pde_solver', 'weno_solver'} <= syms


def test_no_link_time_cuda_dependency():
    """libcuda / nvrtc / nvJitLink / nccl are dlopen()ed on first use, so the
    library loads on a CPU-only box."""
    out = subprocess.check_output(['readelf', '-d', lib_path()]).decode()
    needed = re.findall(r'NEEDED.*\[(.*?)\]', out)
    assert not [n for n in needed if re.search(r'cuda|nvrtc|nvJitLink|nccl|torch', n)], needed


def test_toolchain_versions():
    lib = get_cdll()
    v = [ctypes.c_int() for _ in range(4)]
    assert lib.pypde_b200_version(*[ctypes.byref(x) for x in v]) == 0
    assert (v[0].value, v[1].value) >= (12, 9)   # NVRTC that knows compute_100a
    assert (v[2].value, v[3].value) >= (12, 9)   # nvJitLink that ingests its LTO-IR


@pytest.mark.parametrize('system,ndim,N', [('euler', 1, 2), ('euler', 2, 3), ('advect_nc', 2, 2),
                                           ('navier_stokes', 3, 2), ('reactive_euler', 2, 3)])
def test_jit_builds_sm100a_cubin_without_gpu(system, ndim, N):
    lib = get_cdll()
    F, B, S, V = cuda_sources(system, ndim)
    n = ctypes.c_size_t()
    buf = ctypes.create_string_buffer(8 << 20)
    rc = lib.pypde_b200_compile(F.pointer if F else None, B.pointer if B else None,
                                S.pointer if S else None, ndim, N, V, 0, 0,
                                int(getattr(F, 'second_order', False)), ctypes.byref(n), buf,
                                ctypes.c_size_t(8 << 20))
    assert rc == 0, last_error()
    cubin = buf.raw[:n.value]
    assert cubin[:4] == b'\x7fELF'
    path = '/tmp/pypde_b200_test_%s_%d_%d.cubin' % (system, ndim, N)
    open(path, 'wb').write(cubin)
    out = subprocess.check_output(['cuobjdump', '--dump-resource-usage', path]).decode()
    elf = subprocess.check_output(['cuobjdump', '-elf', path]).decode()
    assert 'sm_100' in elf or 'SM100' in elf.upper()
    for k in ['k_boundaries', 'k_weno_sweep', 'k_cfl', 'k_dt', 'k_dg', 'k_wavespeeds', 'k_faces',
              'k_update']:
        assert 'Function %s:' % k in out, k
    # the specialised kernels exist exactly where the host driver looks for them
    # (solver.cpp: Module::Module): the node-thread predictor without gradient terms in the
    # predictor, the fused Rusanov face kernel, the TMA-fed WENO tile kernels in 2-D and 3-D
    first_order_no_B = B is None and not getattr(F, 'second_order', False)
    assert ('Function k_dg_n:' in out) == (first_order_no_B and N >= 2 and N**ndim <= 32)
    assert ('Function k_dg_g:' in out) == (not first_order_no_B and N >= 2 and N**ndim <= 32)
    assert 'Function k_faces_fused:' in out
    assert ('Function k_weno2d:' in out) == (ndim == 2)
    assert ('Function k_weno3d:' in out) == (ndim == 3)
    if ndim >= 2:
        sass = subprocess.check_output(['cuobjdump', '-sass', '-fun', 'k_weno%dd' % ndim,
                                        path]).decode()
        assert 'UTMALDG' in sass and 'SYNCS' in sass      # TMA bulk-tensor load + mbarrier


def test_jit_links_the_opt_in_wavespeed_function():
    """pypde_b200_set_wavespeed: configurations built while a user_L is set link it in place
    of the eigen-solves (no cold QR code left in the fused face kernel); clearing restores
    the reference's definition."""
    from pypde_b200.systems import euler_wavespeed
    lib = get_cdll()
    F, B, S, V = cuda_sources('euler', 2)
    L = euler_wavespeed(2)

    def build():
        n = ctypes.c_size_t()
        buf = ctypes.create_string_buffer(8 << 20)
        rc = lib.pypde_b200_compile(F.pointer, None, None, 2, 3, V, 0, 0, 0, ctypes.byref(n), buf,
                                    ctypes.c_size_t(8 << 20))
        assert rc == 0, last_error()
        path = '/tmp/pypde_b200_test_ws.cubin'
        open(path, 'wb').write(buf.raw[:n.value])
        sass = subprocess.check_output(['cuobjdump', '-sass', '-fun', 'k_faces_fused',
                                        path]).decode()
        return sass.count('\n')

    assert lib.pypde_b200_set_wavespeed(L.pointer) == 0
    try:
        with_l = build()
    finally:
        assert lib.pypde_b200_set_wavespeed(None) == 0
    without = build()
    assert with_l < without / 4          # the eigen-solver is gone from the kernel


def test_jit_reports_user_function_errors():
    lib = get_cdll()
    bad = CudaSource('extern "C" __device__ void user_F(double* o, const double* q, '
                     'const double* dq, int d) { o[0] = undefined_symbol; }', 'bad_F')
    n = ctypes.c_size_t()
    rc = lib.pypde_b200_compile(bad.pointer, None, None, 1, 2, 1, 0, 0, 0, ctypes.byref(n), None,
                                ctypes.c_size_t(0))
    assert rc != 0
    assert 'undefined_symbol' in last_error()
    missing = CudaSource('extern "C" __device__ void not_user_F() {}', 'missing_F')
    rc = lib.pypde_b200_compile(missing.pointer, None, None, 1, 2, 1, 0, 0, 0, ctypes.byref(n),
                                None, ctypes.c_size_t(0))
    assert rc != 0 and 'user_F' in last_error()


def _has_gpu():
    try:
        ctypes.CDLL('libcuda.so.1')
    except OSError:
        return False
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason='checks the no-GPU failure path')
def test_compute_fails_loudly_without_gpu():
    F, B, S, V = cuda_sources('euler', 1)
    Q0 = np.ones((16, V))
    with pytest.raises(RuntimeError, match='pde_solver failed'):
        pypde_b200.pde_solver(Q0, 0.1, [1.], F=F, order=2, ndt=1, stiff=False)
    with pytest.raises(RuntimeError, match='weno_solver failed'):
        pypde_b200.weno_solver(np.ones((8, 1)), 2)


def test_missing_library_is_an_error(monkeypatch):
    import pypde_b200.utils as U
    monkeypatch.setattr(U, '_LIB', None)
    monkeypatch.setattr(U, 'lib_path', lambda: '/nonexistent/libpypde.so')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        U.get_cdll()


def test_cubin_cache_directory_must_be_private(tmp_path):
    """The on-disk cubin cache is loaded as GPU code: a directory that group / others can
    access, or a symlink, is not used (the link still succeeds, nothing is written)."""
    script = '''
import ctypes, os, sys
from pypde_b200.systems import cuda_sources
from pypde_b200.utils import get_cdll, last_error
F, B, S, V = cuda_sources('burgers', 1)
n = ctypes.c_size_t()
rc = get_cdll().pypde_b200_compile(F.pointer, None, None, 1, 2, V, 0, 0, 0, ctypes.byref(n), None,
                                   ctypes.c_size_t(0))
assert rc == 0, last_error()
print(len(os.listdir(os.environ['PYPDE_B200_CACHE_REAL'])))
'''
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(cache, real, mode):
        os.chmod(real, mode)
        env = dict(os.environ, PYPDE_B200_CACHE=str(cache), PYPDE_B200_CACHE_REAL=str(real),
                   PYTHONPATH=root)
        return int(subprocess.check_output([sys.executable, '-c', script], env=env).split()[-1])

    open_dir = tmp_path / 'open'
    open_dir.mkdir()
    assert run(open_dir, open_dir, 0o777) == 0          # world-writable: refused
    assert run(open_dir, open_dir, 0o750) == 0          # group-readable: refused
    link = tmp_path / 'link'
    private = tmp_path / 'private'
    private.mkdir()
    os.symlink(private, link)
    assert run(link, private, 0o700) == 0               # a symlink: refused
    assert run(private, private, 0o700) == 1            # a private directory: used


def test_shipped_cubins_are_found_through_a_symlink_to_the_library(tmp_path):
    """INTEGRATION.md 1 places libpypde.so into the reference package as a symlink
    (pypde/build/libpypde.so).  Loaded through it — and by soname every later load in that
    process resolves to the same object — the library must still look for the cubins shipped
    next to its REAL location (<dir of the real file>/cubin_cache), not next to the link."""
    import sys
    from pypde_b200.utils import lib_path
    real_dir = os.path.join(os.path.dirname(os.path.realpath(lib_path())), 'cubin_cache')
    if not os.path.isdir(real_dir):
        pytest.skip('no shipped cubin cache (filled by __graft_entry__.build())')
    os.chmod(real_dir, 0o700)
    (tmp_path / 'pypde' / 'build').mkdir(parents=True)
    link = tmp_path / 'pypde' / 'build' / 'libpypde.so'
    os.symlink(lib_path(), link)
    script = '''
import ctypes, os, sys
ctypes.CDLL(%r)
from pypde_b200.systems import cuda_sources
from pypde_b200.utils import get_cdll, last_error
F, B, S, V = cuda_sources('burgers', 1)
n = ctypes.c_size_t()
rc = get_cdll().pypde_b200_compile(F.pointer, None, None, 1, 2, V, 0, 0, 0, ctypes.byref(n), None,
                                   ctypes.c_size_t(0))
assert rc == 0, last_error()
''' % str(link)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    empty = tmp_path / 'user_cache'
    empty.mkdir(mode=0o700)
    env = dict(os.environ, PYPDE_B200_CACHE=str(empty), PYPDE_B200_CACHE_DEBUG='1', PYTHONPATH=root)
    out = subprocess.run([sys.executable, '-c', script], env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "shipped cache '%s'" % real_dir in out.stderr, out.stderr
