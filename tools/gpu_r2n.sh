#!/bin/bash
# round 2, GPU session N (1 GPU): occupancy of the n > 5 wave-speed kernel (GPR, 512^2 and 256^2)
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
timeout 600 python tools/variant_sweep.py occ c4 512 3 2>&1 | tee $O/r2n_occ_sweep_512.txt
timeout 600 python tools/variant_sweep.py occ c4 256 3 2>&1 | tee $O/r2n_occ_sweep_256.txt
du -sm $O
