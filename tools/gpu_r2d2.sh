#!/bin/bash
# round 2, GPU session D2 (8 GPUs of one box): C4 strong scaling, C5 at 256^3 on 8 GPUs,
# C2 weak scaling with the slab bit-identity checks, all through bench.py under torchrun
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > $O/r2m_gpus.txt
run() { # run N port config [extra args...]
  local n=$1 port=$2 cfg=$3; shift 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $n --config $cfg --no-cpu-baseline "$@" \
    > $O/r2m_${cfg}_${n}gpu.json 2> $O/r2m_${cfg}_${n}gpu.err
  python - <<PY || tail -5 $O/r2m_${cfg}_${n}gpu.err
import json
d = json.load(open('$O/r2m_${cfg}_${n}gpu.json'))
print('$cfg x$n', '%.3e' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'],
      'slab_bit_identical', d.get('slab_bit_identical'),
      {k: v['bit_identical'] for k, v in (d.get('slab_checks') or {}).items()})
PY
}
# C4 on 2 and 4 GPUs side by side on disjoint devices (device-timed values; e2e shares the host)
( CUDA_VISIBLE_DEVICES=0,1 run 2 29511 c4 --slab-checks gpr_stiff ) &
( CUDA_VISIBLE_DEVICES=2,3,4,5 run 4 29512 c4 --slab-checks gpr_stiff ) &
wait
run 8 29513 c4 --slab-checks gpr_stiff,advect_nc_BS
run 8 29514 c5 --slab-checks ns3d_second_order
run 8 29515 c2
run 2 29516 c2 --slab-checks euler2d,euler2d_periodic,ns3d_second_order,advect_nc_BS,gpr_stiff
du -sm $O
