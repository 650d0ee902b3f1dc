#!/bin/bash
# round 2, last GPU session (1 GPU), no cache environment (the driver's situation): GPU tests,
# smoke(), the Osher sweep, default bench line and C3 with the exact cluster term of the
# projector form
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider ) > /tmp/pytest_full.log 2>&1
grep -E "passed|failed|real" /tmp/pytest_full.log | tee $O/r2final2_pytest.txt
grep -E "^(FAILED|ERROR)" /tmp/pytest_full.log | head
grep -E "GPU vs reference" /tmp/pytest_full.log > $O/r2final2_parity_lines.txt
grep -E "osher|roe|c3_" $O/r2final2_parity_lines.txt
echo "cubins linked on this box: $(ls ~/.cache/pypde_b200 2>/dev/null | wc -l)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/r2final2_smoke.txt
timeout 600 python tools/variant_sweep.py osher c3 512 3 2>&1 | tee $O/r2final2_osher_sweep.txt
timeout 600 python bench.py > $O/r2final2_bench_c2.json 2> $O/r2final2_bench_c2.err
python -c "import json; d=json.loads([l for l in open('$O/r2final2_bench_c2.json') if l.startswith('{')][-1]); print('c2', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 900 python bench.py --config c3 > $O/r2final2_bench_c3.json 2> $O/r2final2_bench_c3.err
python -c "import json; d=json.loads([l for l in open('$O/r2final2_bench_c3.json') if l.startswith('{')][-1]); print('c3', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernels_ms_per_step'])"
du -sm $O
