#!/bin/bash
# round 2, GPU session M (1 GPU): n > 5 eigen path — permutation step on bit masks, and how many
# local-memory matrices may be in flight per SM (eig sweep on GPR, 256^2 and 512^2)
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
timeout 600 python tools/variant_sweep.py eig c4 256 3 2>&1 | tee $O/r2m_eig_sweep_256.txt
timeout 600 python tools/variant_sweep.py eig c4 512 3 2>&1 | tee $O/r2m_eig_sweep_512.txt
du -sm $O
