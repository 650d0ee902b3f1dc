#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s2e_pytest.log
tools/variant_bench.sh "PDE_NOP=0" "PDE_FACES_SIDE_LOOP=0" > gpurun_out/s2e_variants.log 2>&1
cat gpurun_out/s2e_pytest.log gpurun_out/s2e_variants.log
