#!/bin/bash
# round 2, GPU session S (1 GPU): the projector form of the Osher / Roe dissipation — GPU tests,
# the osher sweep at C3 512^2, ncu of the C3 kernels with the hot lines of k_faces, and the bench
# lines of every configuration with the final kernels
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
T=/tmp/ncu_reps; mkdir -p $T
timeout 900 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $T/pytest_full.log 2>&1
grep -E "passed|failed|error" $T/pytest_full.log | tail -3 | tee $O/r2s_pytest.txt
grep -E "^(FAILED|ERROR)" $T/pytest_full.log | head -20 | tee -a $O/r2s_pytest.txt
grep -E "GPU vs reference" $T/pytest_full.log > $O/r2s_parity_lines.txt
timeout 600 python tools/variant_sweep.py osher c3 512 3 2>&1 | tee $O/r2s_osher_sweep.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:'k_dg_stiff|k_faces' -s 21 -c 3 -o $T/c3 python tools/prof_config.py c3 8 > $O/r2s_ncu.log 2>&1
grep -E "Report|ERROR|rror" $O/r2s_ncu.log | tail -3
cp profiles/r2_ncu_kernels.json $O/r2s_ncu_kernels.json
python tools/ncu_summary.py $T/c3.ncu-rep --into $O/r2s_ncu_kernels.json --config c3 > $O/r2s_ncu_c3.txt 2>&1
python tools/hot_config.py $T/c3.ncu-rep c3 k_faces 45 > $O/r2s_hot_c3_faces.txt 2>&1
cp $O/r2s_ncu_kernels.json profiles/r2_ncu_kernels.json
for c in c2 c1 c3 c4 c5; do
  timeout 900 python bench.py --config $c > $O/r2s_bench_$c.json 2> $O/r2s_bench_$c.err
  python -c "import json,sys; d=json.loads([l for l in open('$O/r2s_bench_$c.json') if l.startswith('{')][-1]); r=d['roofline']; print('$c', '%.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], r['kernel'], r['bound'], 'frac %.3f'%r['frac'], {k:round(v,3) for k,v in r['kernels_ms_per_step'].items()}, 'cpu', d['cpu_baseline']['value'])" || tail -3 $O/r2s_bench_$c.err
done
PYPDE_B200_QUIET=1 timeout 900 python tools/parity_report.py 2>&1 | grep -v "^t = \|^Using\|^compiling" > $O/r2s_parity.txt; grep -c . $O/r2s_parity.txt; grep ABOVE $O/r2s_parity.txt
du -sm $O
