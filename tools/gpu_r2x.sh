#!/bin/bash
# round 2, GPU session X (1 GPU): final tree after the k_dg_stiff Krylov-loop work — GPU tests,
# smoke(), ncu of the C3 kernels with the hot lines of k_dg_stiff, bench lines of C3, C4 and the
# default C2 line
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
T=/tmp/ncu_reps; mkdir -p $T
timeout 900 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $T/pytest_full.log 2>&1
grep -E "passed|failed|error" $T/pytest_full.log | tail -3 | tee $O/r2x_pytest.txt
grep -E "^(FAILED|ERROR)" $T/pytest_full.log | head -20 | tee -a $O/r2x_pytest.txt
grep -E "GPU vs reference" $T/pytest_full.log > $O/r2x_parity_lines.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/r2x_smoke.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:'k_dg_stiff|k_faces' -s 21 -c 3 -o $T/c3 python tools/prof_config.py c3 8 > $O/r2x_ncu.log 2>&1
grep -E "Report|ERROR|rror" $O/r2x_ncu.log | tail -3
cp profiles/r2_ncu_kernels.json $O/r2x_ncu_kernels.json
python tools/ncu_summary.py $T/c3.ncu-rep --into $O/r2x_ncu_kernels.json --config c3 > $O/r2x_ncu_c3.txt 2>&1
python tools/hot_config.py $T/c3.ncu-rep c3 k_dg_stiff 45 > $O/r2x_hot_c3_dg_stiff.txt 2>&1
cp $O/r2x_ncu_kernels.json profiles/r2_ncu_kernels.json
for c in c3 c4 c2; do
  timeout 900 python bench.py --config $c > $O/r2x_bench_$c.json 2> $O/r2x_bench_$c.err
  python -c "import json,sys; d=json.loads([l for l in open('$O/r2x_bench_$c.json') if l.startswith('{')][-1]); r=d['roofline']; print('$c', '%.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], r['kernel'], r['bound'], 'frac %.3f'%r['frac'], {k:round(v,3) for k,v in r['kernels_ms_per_step'].items()})" || tail -3 $O/r2x_bench_$c.err
done
PYPDE_B200_QUIET=1 timeout 900 python tools/parity_report.py 2>&1 | grep -v "^t = \|^Using\|^compiling" > $O/r2x_parity.txt; grep -c . $O/r2x_parity.txt; grep ABOVE $O/r2x_parity.txt
du -sm $O
