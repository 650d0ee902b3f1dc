#!/bin/bash
tools/variant_bench.sh "PDE_NOP=0" "PDE_EIG_LOCKSTEP=0" "PDE_DGN_MINBLOCKS=12" > gpurun_out/s2d_variants.log 2>&1
tools/variant_env.sh "PYPDE_B200_WS_BLOCK=128 PYPDE_B200_WS_MINBLOCKS=5" "PYPDE_B200_WS_BLOCK=192 PYPDE_B200_WS_MINBLOCKS=3" "PYPDE_B200_WS_BLOCK=256 PYPDE_B200_WS_MINBLOCKS=3" "PYPDE_B200_WS_BLOCK=128 PYPDE_B200_WS_MINBLOCKS=3" >> gpurun_out/s2d_variants.log 2>&1
cat gpurun_out/s2d_variants.log
