#!/bin/bash
# round 2, GPU session J (1 GPU): the GPU test-suite with the final kernels, the default bench
# line (now with the smooth-periodic variant beside the explosion), and the last two sweeps:
# C5 (viscous warm start, DG minblocks, k_faces_side block shapes) and the GPR stiff occupancy
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 | tee $O/r2j_pytest.txt
timeout 600 python bench.py > $O/r2j_bench_c2.json 2> $O/r2j_bench_c2.err; tail -c 600 $O/r2j_bench_c2.json
timeout 600 python tools/variant_sweep.py c5 c5 128 3 2>&1 | tee $O/r2j_c5_sweep.txt
timeout 600 python tools/variant_sweep.py gprstiff c4 512 3 2>&1 | tee $O/r2j_gprstiff_sweep.txt
du -sm $O
