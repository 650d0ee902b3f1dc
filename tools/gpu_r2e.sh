#!/bin/bash
# round 2, GPU session E (1 GPU): k_dg_g / second-order k_faces_side parity and sweeps
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > /tmp/pytest_full.log 2>&1
grep -E "passed|failed|error" /tmp/pytest_full.log | tail -3
grep -E "^(FAILED|ERROR|SKIPPED)" /tmp/pytest_full.log | head -20
tail -160 /tmp/pytest_full.log > $O/r2e_pytest.log
python tools/variant_sweep.py c5 c5 128 3 > $O/r2e_c5_sweep.log 2>&1; cat $O/r2e_c5_sweep.log
python tools/variant_sweep.py eig c4 256 3 > $O/r2e_eig_sweep.log 2>&1; cat $O/r2e_eig_sweep.log
for c in c4 c5; do
  python bench.py --config $c --no-cpu-baseline > $O/r2e_bench_$c.json 2> $O/r2e_bench_$c.err
  python -c "import json,sys; d=json.load(open('$O/r2e_bench_$c.json')); print('$c', '%.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel'], {k:round(v,3) for k,v in d['roofline']['kernels_ms_per_step'].items()})" || tail -3 $O/r2e_bench_$c.err
done
du -sm $O
