#!/bin/bash
for cfg in "256 0" "288 112" "192 112" "576 112"; do
  set -- $cfg
  echo "=== FF_BLOCK=$1 MAXNREG=$2"
  PYPDE_B200_FF_BLOCK=$1 PYPDE_B200_EXTRA_DEFINES="PDE_FF_MAXNREG=$2" python bench.py --no-e2e --no-cpu-baseline --steps 6 --warmup 6 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_ms_per_step']; print('%.3e cu/s  %.2f ms/step  '%(d['value'], d['ms_per_step'])+'  '.join('%s=%.2f'%(a.replace('k_',''),b) for a,b in k.items() if b>0.3))"
done > gpurun_out/s2p_variants.log 2>&1
cat gpurun_out/s2p_variants.log
