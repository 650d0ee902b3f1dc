#!/bin/bash
# round 2, GPU session A: parity at the survey's sizes, all-config bench, stiff-kernel sweep, ncu
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/r2a_gpu.txt
( time python -m pytest tests -m gpu -q -rA 2>&1 | tail -150 ) > $O/r2a_pytest.log 2>&1
tail -3 $O/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2a_smoke.log 2>&1; tail -1 $O/r2a_smoke.log
python tools/parity_report.py > $O/r2a_parity.txt 2>&1
for c in c2 c1 c2smooth c3 c4 c5; do
  python bench.py --config $c > $O/r2a_bench_$c.json 2> $O/r2a_bench_$c.err
  python -c "import json,sys; d=json.load(open('$O/r2a_bench_$c.json')); print('$c', '%.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['kernel'], {k:round(v,3) for k,v in d['roofline']['kernels_ms_per_step'].items()}, 'cpu', d['cpu_baseline']['value'])" || tail -3 $O/r2a_bench_$c.err
done
for uf in numba traced; do
  python bench.py --config c2 --user-functions $uf --no-cpu-baseline > $O/r2a_bench_c2_$uf.json 2> $O/r2a_bench_c2_$uf.err
  python -c "import json; d=json.load(open('$O/r2a_bench_c2_$uf.json')); print('c2 $uf', '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])" || tail -3 $O/r2a_bench_c2_$uf.err
done
python tools/variant_sweep.py stiff c3 512 4 > $O/r2a_stiff_sweep.log 2>&1; cat $O/r2a_stiff_sweep.log
python tools/variant_sweep.py eig c4 256 4 > $O/r2a_eig_sweep.log 2>&1; cat $O/r2a_eig_sweep.log
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:k_faces_side -s 14 -c 2 -o $O/r2a_c2 python tools/prof_config.py c2 8 > $O/r2a_ncu.log 2>&1
$NCU -k regex:'k_dg_stiff|k_faces' -s 21 -c 3 -o $O/r2a_c3 python tools/prof_config.py c3 8 >> $O/r2a_ncu.log 2>&1
$NCU -k regex:'k_wavespeeds|k_dg_stiff|k_faces|k_cfl' -s 28 -c 5 -o $O/r2a_c4 python tools/prof_config.py c4 8 >> $O/r2a_ncu.log 2>&1
$NCU -k regex:'k_wavespeeds|k_dg|k_faces|k_weno_sweep|k_cfl' -s 70 -c 10 -o $O/r2a_c5 python tools/prof_config.py c5 8 >> $O/r2a_ncu.log 2>&1
tail -5 $O/r2a_ncu.log
ls -la $O/*.ncu-rep
