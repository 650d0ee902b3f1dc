#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s2k_pytest.log
tools/variant_bench.sh "PDE_NOP=0" > gpurun_out/s2k_variants.log 2>&1
python tools/config_survey.py big 2>&1 | grep -v "^t = \|Using\|compiling" > gpurun_out/s2k_survey.log
cat gpurun_out/s2k_pytest.log gpurun_out/s2k_variants.log gpurun_out/s2k_survey.log
