#!/bin/bash
# round 2, GPU session R (1 GPU): the n > 5 eigen path as it now stands (two-pass Jacobian, bit-mask
# permutation step, certified characteristic polynomial) — GPU tests, ncu of the GPR kernels with
# the hot lines of k_wavespeeds, the C4 bench line
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
T=/tmp/ncu_reps; mkdir -p $T
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 | tee $O/r2r_pytest.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:'k_wavespeeds|k_dg_stiff|k_faces|k_cfl' -s 35 -c 5 -o $T/c4 python tools/prof_config.py c4 8 > $O/r2r_ncu.log 2>&1
grep -E "Report|ERROR|rror" $O/r2r_ncu.log | tail -3
cp profiles/r2_ncu_kernels.json $O/r2r_ncu_kernels.json
python tools/ncu_summary.py $T/c4.ncu-rep --into $O/r2r_ncu_kernels.json --config c4 > $O/r2r_ncu_c4.txt 2>&1
python tools/hot_config.py $T/c4.ncu-rep c4 k_wavespeeds 60 > $O/r2r_hot_c4_wavespeeds.txt 2>&1
cp $O/r2r_ncu_kernels.json profiles/r2_ncu_kernels.json
timeout 600 python bench.py --config c4 > $O/r2r_bench_c4.json 2> $O/r2r_bench_c4.err; tail -c 1500 $O/r2r_bench_c4.json
timeout 600 python tools/variant_sweep.py eig c4 512 3 2>&1 | tee $O/r2r_eig_sweep_512.txt
timeout 600 python tools/variant_sweep.py eig c4 256 3 2>&1 | tee $O/r2r_eig_sweep_256.txt
du -sm $O
