#!/bin/bash
# ncu --set full of the four main kernels of one step (C2), after the 6 start-up steps
ncu --set full --clock-control none --import-source on -k regex:'k_dg|k_faces_fused|k_weno_sweep|k_cfl' -s 18 -c 6 -o gpurun_out/prof_h python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/s2_ncu.log 2>&1
tail -2 gpurun_out/s2_ncu.log
