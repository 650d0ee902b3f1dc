#!/bin/bash
# usage: tools/variant_env.sh "VAR=val VAR2=val" ...: device-resident bench per environment setting
for envs in "$@"; do
  echo "=== $envs"
  env $envs python bench.py --no-e2e --no-cpu-baseline --steps 6 --warmup 6 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels_ms_per_step']; print('%.3e cu/s  %.2f ms/step  '%(d['value'], d['ms_per_step'])+'  '.join('%s=%.2f'%(a.replace('k_',''),b) for a,b in k.items() if b>0.3))"
done
