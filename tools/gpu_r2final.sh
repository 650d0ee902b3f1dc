#!/bin/bash
# round 2, final GPU session (1 GPU): exactly what the driver runs at round end, with NO cache
# environment (the cubins shipped next to libpypde.so must be found on their own) — GPU tests,
# smoke(), the default bench line; then the C3 / C4 lines of the final tree
O=gpurun_out
mkdir -p $O
ls pypde_b200/build/cubin_cache | wc -l
( time timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) 2>&1 | tail -6 | tee $O/r2final_pytest.txt
ls ~/.cache/pypde_b200 2>/dev/null | wc -l
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/r2final_smoke.txt
timeout 600 python bench.py > $O/r2final_bench_c2.json 2> $O/r2final_bench_c2.err; tail -c 300 $O/r2final_bench_c2.json
for c in c3 c4; do
  timeout 900 python bench.py --config $c > $O/r2final_bench_$c.json 2> $O/r2final_bench_$c.err
  python -c "import json,sys; d=json.loads([l for l in open('$O/r2final_bench_$c.json') if l.startswith('{')][-1]); r=d['roofline']; print('$c', '%.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], r['kernel'], 'frac %.3f'%r['frac'], {k:round(v,3) for k,v in r['kernels_ms_per_step'].items()})" || tail -3 $O/r2final_bench_$c.err
done
du -sm $O
