#!/bin/bash
# round 2, GPU session B: full parity log, all-config bench lines, kernel sweeps, ncu summaries
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/r2b_pytest_full.log 2>&1
grep -E "passed|failed|error" $O/r2b_pytest_full.log | tail -3
grep -E "^(FAILED|ERROR)" $O/r2b_pytest_full.log | head -20
grep -E "GPU vs reference" $O/r2b_pytest_full.log > $O/r2b_sized_parity.txt
tail -400 $O/r2b_pytest_full.log > $O/r2b_pytest.log; rm $O/r2b_pytest_full.log
python tools/parity_report.py > $O/r2b_parity.txt 2>&1; grep -c . $O/r2b_parity.txt; grep ABOVE $O/r2b_parity.txt
for c in c2 c1 c2smooth c3 c4 c5; do
  python bench.py --config $c > $O/r2b_bench_$c.json 2> $O/r2b_bench_$c.err
  python -c "import json,sys; d=json.load(open('$O/r2b_bench_$c.json')); print('$c', '%.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel'], {k:round(v,3) for k,v in d['roofline']['kernels_ms_per_step'].items()}, 'cpu', d['cpu_baseline']['value'])" || tail -3 $O/r2b_bench_$c.err
done
for uf in numba traced; do
  python bench.py --config c2 --user-functions $uf --no-cpu-baseline > $O/r2b_bench_c2_$uf.json 2> $O/r2b_bench_c2_$uf.err
  python -c "import json; d=json.load(open('$O/r2b_bench_c2_$uf.json')); print('c2 $uf', '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])" || tail -3 $O/r2b_bench_c2_$uf.err
done
python tools/variant_sweep.py faces c2 2048 4 > $O/r2b_faces_sweep.log 2>&1; cat $O/r2b_faces_sweep.log
python tools/variant_sweep.py weno3d c5 128 3 > $O/r2b_weno3d_sweep.log 2>&1; cat $O/r2b_weno3d_sweep.log
python tools/variant_sweep.py stiff c3 512 3 > $O/r2b_stiff_sweep.log 2>&1; cat $O/r2b_stiff_sweep.log
python tools/variant_sweep.py eig c4 256 3 > $O/r2b_eig_sweep.log 2>&1; cat $O/r2b_eig_sweep.log
NCU="ncu --set full --clock-control none --import-source on -f"
# one launch of each kernel of interest, a full-dt step (launch index = steps x launches per step)
$NCU -k regex:'k_faces_side|k_dg_n|k_weno2d|k_cfl_q' -s 24 -c 5 -o $O/r2b_c2 python tools/prof_config.py c2 8 > $O/r2b_ncu.log 2>&1
$NCU -k regex:'k_dg_stiff|k_faces' -s 14 -c 3 -o $O/r2b_c3 python tools/prof_config.py c3 8 512 >> $O/r2b_ncu.log 2>&1
$NCU -k regex:'k_wavespeeds|k_dg_stiff|k_faces|k_cfl' -s 35 -c 5 -o $O/r2b_c4 python tools/prof_config.py c4 8 >> $O/r2b_ncu.log 2>&1
$NCU -k regex:'k_wavespeeds|k_dg|k_faces|k_weno3d|k_cfl' -s 49 -c 7 -o $O/r2b_c5 python tools/prof_config.py c5 8 128 >> $O/r2b_ncu.log 2>&1
grep -E "Report|ERROR|rror" $O/r2b_ncu.log | tail -8
for c in c2 c3 c4 c5; do
  python tools/ncu_summary.py $O/r2b_$c.ncu-rep --into $O/r2b_ncu_kernels.json --config $c > $O/r2b_ncu_$c.txt 2>&1
done
grep -E "^==|gpu__time_duration|pipe_fp64|FP64 flops|top stall" $O/r2b_ncu_c*.txt | head -80
ls -la $O/*.ncu-rep
# the reports are large: keep what fits the 64 MiB return limit (the stiff kernel's first)
rm -f $O/r2b_c4.ncu-rep
du -sm $O
