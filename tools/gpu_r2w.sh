#!/bin/bash
# round 2, GPU session W (1 GPU): the Gram-Schmidt loop of k_dg_stiff (prefetch / two basis vectors
# per pass) at C3 512^2 and GPR 512^2
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
timeout 600 python tools/variant_sweep.py stiff2 c3 512 3 2>&1 | tee $O/r2w_stiff2_c3.txt
timeout 600 python tools/variant_sweep.py stiff2 c4 512 3 2>&1 | tee $O/r2w_stiff2_c4.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -s -k 'stiff or c3_ or c4_' 2>&1 | grep -E 'GPU vs reference|passed|failed' | tee $O/r2w_stiff_parity.txt
du -sm $O
