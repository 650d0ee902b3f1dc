#!/bin/bash
python tools/config_survey.py big C4,C5 2>&1 | grep -v "^t = \|Using\|compiling" > gpurun_out/s2j.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_wavespeeds' -s 2 -c 1 -o gpurun_out/prof_c5 python tools/prof_c5.py > gpurun_out/s2j_ncu.log 2>&1
cat gpurun_out/s2j.log; tail -2 gpurun_out/s2j_ncu.log
