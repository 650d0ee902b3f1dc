#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s2b_pytest.log
tools/variant_bench.sh "PDE_NOP=0" "PDE_DG_MINBLOCKS=24" "PDE_DG_MINBLOCKS=32" > gpurun_out/s2b_variants.log 2>&1
tools/variant_env.sh "PYPDE_B200_DG_CARVEOUT=-1" "PYPDE_B200_DG_CARVEOUT=50" >> gpurun_out/s2b_variants.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_dg|k_faces_fused|k_weno_sweep|k_cfl' -s 18 -c 6 -o gpurun_out/prof_g python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/s2b_ncu.log 2>&1
cat gpurun_out/s2b_pytest.log gpurun_out/s2b_variants.log; tail -3 gpurun_out/s2b_ncu.log
