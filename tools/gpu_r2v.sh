#!/bin/bash
# round 2, GPU session V (1 GPU): what the driver runs at round end, on the final tree — the GPU
# tests, smoke(), the default bench line and the reference arm
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 | tee $O/r2v_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r2v_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2v_bench_c2_k20.json 2> $O/r2v_bench_c2_k20.err; tail -c 700 $O/r2v_bench_c2_k20.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/r2v_bench_c2_reference.json 2> $O/r2v_bench_c2_reference.err; tail -c 500 $O/r2v_bench_c2_reference.json
timeout 600 python bench.py --config c3 --no-cpu-baseline > $O/r2v_bench_c3.json 2> $O/r2v_bench_c3.err; python -c "import json; d=json.loads([l for l in open('$O/r2v_bench_c3.json') if l.startswith('{')][-1]); print('c3', d['value'], d['ms_per_step'])"
du -sm $O
