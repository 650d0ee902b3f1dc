#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:'k_dg_stiff' -s 2 -c 1 -o gpurun_out/prof_c3 python tools/prof_c3.py > gpurun_out/s2q_ncu.log 2>&1
tail -3 gpurun_out/s2q_ncu.log
