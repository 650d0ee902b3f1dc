#!/bin/bash
# round 2, GPU session G (1 GPU): shared-memory workspace of the V > 5 eigen-iteration
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > /tmp/pytest_full.log 2>&1
grep -E "passed|failed|error" /tmp/pytest_full.log | tail -3
grep -E "^(FAILED|ERROR|SKIPPED)" /tmp/pytest_full.log | head -20
tail -170 /tmp/pytest_full.log > $O/r2g_pytest.log
python tools/variant_sweep.py eig c4 256 3 > $O/r2g_eig_sweep.log 2>&1; cat $O/r2g_eig_sweep.log
python tools/variant_sweep.py eig c4 512 3 > $O/r2g_eig_sweep_512.log 2>&1; cat $O/r2g_eig_sweep_512.log
python bench.py --config c4 --no-cpu-baseline > $O/r2g_bench_c4.json 2> $O/r2g_bench_c4.err
python -c "import json,sys; d=json.load(open('$O/r2g_bench_c4.json')); print('c4', '%.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], {k:round(v,3) for k,v in d['roofline']['kernels_ms_per_step'].items()})" || tail -3 $O/r2g_bench_c4.err
PYPDE_B200_QUIET=1 python tools/parity_report.py 2>&1 | grep -E "gpr|c4_" 
du -sm $O
