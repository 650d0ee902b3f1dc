#!/bin/bash
tools/variant_bench.sh "PDE_NOP=0" "PDE_NOP=1" > gpurun_out/s2m_variants.log 2>&1
cat gpurun_out/s2m_variants.log
