#!/bin/bash
# round 2, GPU session Z (1 GPU): fused multiply-adds in the Krylov linear algebra of k_dg_stiff
# (C3 512^2 sweep; parity of the stiff cases)
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
timeout 900 python tools/variant_sweep.py stiff2 c3 512 3 2>&1 | tee $O/r2z_stiff2_c3.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -s -k 'stiff or c3_ or c4_' 2>&1 | grep -E 'GPU vs reference|passed|failed' | tee $O/r2z_stiff_parity.txt
du -sm $O
