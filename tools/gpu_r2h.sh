#!/bin/bash
# round 2, GPU session H (1 GPU): occupancy against registers for the QR-bound kernels
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > /tmp/pytest_full.log 2>&1
grep -E "passed|failed|error" /tmp/pytest_full.log | tail -3
grep -E "^(FAILED|ERROR)" /tmp/pytest_full.log | head -20
python tools/variant_sweep.py occ c4 256 3 > $O/r2h_occ_c4.log 2>&1; cat $O/r2h_occ_c4.log
python tools/variant_sweep.py occ c3 512 3 > $O/r2h_occ_c3.log 2>&1; cat $O/r2h_occ_c3.log
du -sm $O
