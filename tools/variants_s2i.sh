#!/bin/bash
for e in "PYPDE_B200_WS_BLOCK=256 PYPDE_B200_WS_MINBLOCKS=2" "PYPDE_B200_WS_BLOCK=512 PYPDE_B200_WS_MINBLOCKS=1" "PYPDE_B200_WS_BLOCK=128 PYPDE_B200_WS_MINBLOCKS=2"; do
  echo "=== $e"
  env $e python tools/config_survey.py big C4,C5 2>&1 | grep -v "^t = \|Using\|compiling"
done > gpurun_out/s2i.log 2>&1
cat gpurun_out/s2i.log
