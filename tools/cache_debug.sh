#!/bin/bash
# the driver's situation: no PYPDE_B200_CACHE; the GPU tests must find the cubins shipped next to
# the library (also after tests/test_dropin.py has loaded it through a symlink)
export PYPDE_B200_CACHE_DEBUG=1 PYPDE_B200_QUIET=1
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -p no:cacheprovider -s ) > /tmp/dbg.log 2>&1
echo "cubins linked on this box: $(ls ~/.cache/pypde_b200 2>/dev/null | wc -l)"
grep -E "passed|failed|real" /tmp/dbg.log
