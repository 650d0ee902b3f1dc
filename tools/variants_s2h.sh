#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s2h_pytest.log
tools/variant_env.sh "PYPDE_B200_WENO_FUSED=1" "PYPDE_B200_WENO_FUSED=0" > gpurun_out/s2h_variants.log 2>&1
python tools/config_survey.py big 2>&1 | grep -v "^t = \|Using\|compiling" > gpurun_out/s2h_survey.log
cat gpurun_out/s2h_pytest.log gpurun_out/s2h_variants.log gpurun_out/s2h_survey.log
