#!/bin/bash
# end-of-round evidence run on one B200 (gpurun -- tools/evidence_run.sh): parity suite, smoke, both bench arms,
# launch list, ncu --set full of the main kernels, parity table, config survey; outputs under gpurun_out/
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > gpurun_out/f_smoke.log
python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/f_bench_ref.json
python bench.py 2>&1 | tail -1 > gpurun_out/f_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f_launches.csv python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/f_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_dg|k_faces_fused|k_weno2d|k_cfl' -s 15 -c 5 -o gpurun_out/prof_i python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/f_ncu.log 2>&1
cat gpurun_out/f_pytest.log gpurun_out/f_smoke.log; cut -c1-400 gpurun_out/f_bench_ref.json; cut -c1-1200 gpurun_out/f_bench.json; tail -2 gpurun_out/f_ncu.log
python tools/parity_report.py 2>&1 | grep -v "^t = \|Using\|compiling" > gpurun_out/f_parity.txt
python tools/config_survey.py big 2>&1 | grep -v "^t = \|Using\|compiling" > gpurun_out/f_survey.log
tail -3 gpurun_out/f_parity.txt; cat gpurun_out/f_survey.log
