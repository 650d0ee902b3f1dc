#!/bin/bash
python -m pytest tests -m gpu -x -q -k "solver_golden or variants" 2>&1 | tail -5 > gpurun_out/s2l_pytest.log
tools/variant_bench.sh "PDE_NOP=0" > gpurun_out/s2l_variants.log 2>&1
python tools/config_survey.py big C2,C5 2>&1 | grep -v "^t = \|Using\|compiling" > gpurun_out/s2l_survey.log
cat gpurun_out/s2l_pytest.log gpurun_out/s2l_variants.log gpurun_out/s2l_survey.log
