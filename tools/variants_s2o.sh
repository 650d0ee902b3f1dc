#!/bin/bash
python -m pytest tests -m gpu -x -q -k "variants or solver_golden" 2>&1 | tail -4 > gpurun_out/s2o_pytest.log
tools/variant_bench.sh "PDE_NOP=0" > gpurun_out/s2o_variants.log 2>&1
python tools/config_survey.py big C2 2>&1 | grep -v "^t = \|Using\|compiling" >> gpurun_out/s2o_variants.log
cat gpurun_out/s2o_pytest.log gpurun_out/s2o_variants.log
