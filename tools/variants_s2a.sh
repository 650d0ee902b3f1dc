#!/bin/bash
# one GPU call: parity suite on the new defaults, then A/B of each new switch
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s2a_pytest.log
tools/variant_bench.sh "PDE_NOP=0" "PDE_EIG_HALLEY=0" "PDE_DG_TRACE_BY_SIDE=0" "PDE_DG_PREFETCH=0" "PDE_DG_MINBLOCKS=16" "PDE_DG_MINBLOCKS=24" "PDE_DG_MINBLOCKS=32" "PDE_DG_MINBLOCKS=32;PDE_DG_PREFETCH=0" > gpurun_out/s2a_variants.log 2>&1
tools/variant_env.sh "PYPDE_B200_WS_BLOCK=256 PYPDE_B200_WS_MINBLOCKS=2" "PYPDE_B200_WS_BLOCK=128 PYPDE_B200_WS_MINBLOCKS=4" >> gpurun_out/s2a_variants.log 2>&1
cat gpurun_out/s2a_pytest.log gpurun_out/s2a_variants.log
