#!/bin/bash
# Build box (no GPU): JIT every kernel configuration the GPU test-suite, the benches and
# the sweeps use into pypde_b200/build/cubin_cache (git-ignored; travels to the GPU box with
# the gpurun snapshot), so that GPU sessions start without compiling.
cd "$(dirname "$0")/.."
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache PYPDE_B200_PREBUILD=1
rm -rf $PYPDE_B200_CACHE; mkdir -p $PYPDE_B200_CACHE; chmod 700 $PYPDE_B200_CACHE
python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -1
# (a test that builds several variants stops at its first one in prebuild mode: the variants of
#  test_large_system_eigen_paths one by one)
for d in PDE_EIG_TWOPASS=0 PDE_EIG_MASK=0 PDE_EIG_HESS_POLY=0 PDE_EIG_DEFLATE=0; do
  PYPDE_B200_EXTRA_DEFINES=$d python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider \
    -k "test_solver_golden and gpr" > /dev/null 2>&1
done
PYPDE_B200_EXTRA_DEFINES=PDE_ABS_POLY=0 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider \
  -k "test_solver_golden and (osher or roe)" > /dev/null 2>&1
for c in c1 c2 c3 c4 c5; do python tools/prof_config.py $c 1 16 > /dev/null 2>&1; done
for s in "$@"; do python tools/variant_sweep.py $s x 16 1 2>&1 | grep -c failed; done
ls $PYPDE_B200_CACHE | wc -l
