#!/bin/bash
python tools/config_survey.py big 2>&1 | grep -v "^t = \|Using\|compiling" > gpurun_out/s2g_survey.log
cat gpurun_out/s2g_survey.log
