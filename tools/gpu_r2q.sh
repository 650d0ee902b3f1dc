#!/bin/bash
# round 2, GPU session Q (1 GPU): details of the n > 5 eigen path (row pitch of the active block,
# zero-factor skips in the Hessenberg reduction, balancing sweeps) on GPR at 512^2 and 256^2
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
timeout 600 python tools/variant_sweep.py eig2 c4 512 3 2>&1 | tee $O/r2q_eig2_sweep_512.txt
timeout 600 python tools/variant_sweep.py eig2 c4 256 3 2>&1 | tee $O/r2q_eig2_sweep_256.txt
du -sm $O
