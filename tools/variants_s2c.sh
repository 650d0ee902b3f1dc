#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s2c_pytest.log
tools/variant_bench.sh "PDE_NOP=0" "PDE_DGN_MINBLOCKS=20" "PDE_DGN_MINBLOCKS=24" "PDE_DGN_MINBLOCKS=32" > gpurun_out/s2c_variants.log 2>&1
tools/variant_env.sh "PYPDE_B200_DG_NODE=0" >> gpurun_out/s2c_variants.log 2>&1
cat gpurun_out/s2c_pytest.log gpurun_out/s2c_variants.log
