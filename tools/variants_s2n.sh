#!/bin/bash
tools/variant_bench.sh "PDE_NOP=0" "PDE_EIG_ITER_NOINLINE=0" > gpurun_out/s2n_variants.log 2>&1
cat gpurun_out/s2n_variants.log
