#!/bin/bash
# round 2, GPU session F (1 GPU): the evidence run — parity, ncu summaries first (their JSON
# feeds bench.py's roofline object), bench lines of every configuration, launch list.
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
T=/tmp/ncu_reps; mkdir -p $T
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $T/pytest_full.log 2>&1
grep -E "passed|failed|error" $T/pytest_full.log | tail -3
grep -E "^(FAILED|ERROR|SKIPPED)" $T/pytest_full.log | head -20
grep -E "GPU vs reference" $T/pytest_full.log > $O/r2f_sized_parity.txt
tail -170 $T/pytest_full.log > $O/r2f_pytest.log
PYPDE_B200_QUIET=1 python tools/parity_report.py 2>&1 | grep -v "^t = \|^Using\|^compiling" > $O/r2f_parity.txt; grep -c . $O/r2f_parity.txt; grep ABOVE $O/r2f_parity.txt
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:'k_faces_side|k_dg_n|k_weno2d|k_cfl_q' -s 35 -c 5 -o $T/c2 python tools/prof_config.py c2 8 > $O/r2f_ncu.log 2>&1
$NCU -k regex:'k_faces_side|k_dg_n' -s 21 -c 3 -o $T/c2smooth python tools/prof_config.py c2smooth 8 >> $O/r2f_ncu.log 2>&1
$NCU -k regex:'k_dg_stiff|k_faces' -s 21 -c 3 -o $T/c3 python tools/prof_config.py c3 8 >> $O/r2f_ncu.log 2>&1
$NCU -k regex:'k_wavespeeds|k_dg_stiff|k_faces|k_cfl' -s 35 -c 5 -o $T/c4 python tools/prof_config.py c4 8 >> $O/r2f_ncu.log 2>&1
$NCU -k regex:'k_faces_side|k_dg_g|k_weno_sweep|k_cfl' -s 56 -c 8 -o $T/c5 python tools/prof_config.py c5 8 >> $O/r2f_ncu.log 2>&1
grep -E "Report|ERROR|rror" $O/r2f_ncu.log | tail -8
rm -f $O/r2f_ncu_kernels.json
for c in c2 c2smooth c3 c4 c5; do
  python tools/ncu_summary.py $T/$c.ncu-rep --into $O/r2f_ncu_kernels.json --config $c > $O/r2f_ncu_$c.txt 2>&1
done
cp $O/r2f_ncu_kernels.json profiles/r2_ncu_kernels.json
python tools/hot_config.py $T/c3.ncu-rep c3 k_dg_stiff 45 > $O/r2f_hot_c3_dg_stiff.txt 2>&1
python tools/hot_config.py $T/c5.ncu-rep c5 k_faces_side 35 > $O/r2f_hot_c5_faces_side.txt 2>&1
python tools/hot_config.py $T/c5.ncu-rep c5 k_dg_g 35 > $O/r2f_hot_c5_dg_g.txt 2>&1
python tools/hot_config.py $T/c4.ncu-rep c4 k_wavespeeds 45 > $O/r2f_hot_c4_wavespeeds.txt 2>&1
for c in c2 c1 c2smooth c3 c4 c5; do
  python bench.py --config $c > $O/r2f_bench_$c.json 2> $O/r2f_bench_$c.err
  python -c "import json,sys; d=json.load(open('$O/r2f_bench_$c.json')); r=d['roofline']; print('$c', '%.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], r['kernel'], r['bound'], 'frac %.3f'%r['frac'], {k:round(v,3) for k,v in r['kernels_ms_per_step'].items()}, 'cpu', d['cpu_baseline']['value'])" || tail -3 $O/r2f_bench_$c.err
done
python bench.py --config c2 --steps 20 > $O/r2f_bench_c2_k20.json 2> $O/r2f_bench_c2_k20.err
python bench.py --impl reference --steps 20 --warmup 6 > $O/r2f_bench_c2_reference.json 2> $O/r2f_bench_c2_reference.err
for uf in numba traced; do
  python bench.py --config c2 --user-functions $uf --no-cpu-baseline > $O/r2f_bench_c2_$uf.json 2> $O/r2f_bench_c2_$uf.err
  python -c "import json; d=json.load(open('$O/r2f_bench_c2_$uf.json')); print('c2 $uf', '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])" || tail -3 $O/r2f_bench_c2_$uf.err
done
python tools/variant_sweep.py stiff c3 512 3 > $O/r2f_stiff_sweep.log 2>&1; head -3 $O/r2f_stiff_sweep.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $T/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-smooth > /dev/null 2>&1
python tools/summarize_launches.py $T/launches.csv "ncu launch list of: python bench.py --steps 2 --warmup 3 (not a bench value)" > $O/r2f_launches.txt 2>&1
du -sm $O
