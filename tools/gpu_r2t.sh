#!/bin/bash
# round 2, GPU session T (8 GPUs of one box): C4 strong scaling (512^2 in total on 1 / 2 / 4 / 8
# GPUs) and C3 weak scaling on 8 with the final kernels (the V > 5 eigen path and the projector
# form of the Osher dissipation changed both since session I)
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
run() { # run N port config tag [extra args...]
  local n=$1 port=$2 cfg=$3 tag=$4; shift 4
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $n --config $cfg --no-cpu-baseline "$@" \
    > $O/r2t_${cfg}_${n}gpu$tag.json 2> $O/r2t_${cfg}_${n}gpu$tag.err
  python - <<PY || tail -5 $O/r2t_${cfg}_${n}gpu$tag.err
import json
lines = [l for l in open('$O/r2t_${cfg}_${n}gpu$tag.json') if l.startswith('{')]
d = json.loads(lines[-1])
print('$cfg x$n $tag', '%.3e' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'],
      'slab_bit_identical', d.get('slab_bit_identical'), len(lines), 'json line(s)')
PY
}
( CUDA_VISIBLE_DEVICES=0,1 run 2 29511 c4 "" --slab-checks gpr_stiff ) &
( CUDA_VISIBLE_DEVICES=2,3,4,5 run 4 29512 c4 "" --slab-checks gpr_stiff ) &
( CUDA_VISIBLE_DEVICES=6 python bench.py --config c4 --no-cpu-baseline > $O/r2t_c4_1gpu.json 2> $O/r2t_c4_1gpu.err ) &
wait
run 8 29513 c4 "" --slab-checks gpr_stiff
run 8 29514 c3 "" --slab-checks none
du -sm $O
