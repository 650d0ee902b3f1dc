#!/bin/bash
# round 2, GPU session Y (8 GPUs of one box): C3 weak scaling and C4 strong scaling at 8 ranks on
# the final tree
export PYPDE_B200_CACHE=$PWD/pypde_b200/build/cubin_cache
chmod 700 $PYPDE_B200_CACHE 2>/dev/null
O=gpurun_out
mkdir -p $O
run() { # run N port config tag [extra args...]
  local n=$1 port=$2 cfg=$3 tag=$4; shift 4
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $n --config $cfg --no-cpu-baseline "$@" \
    > $O/r2y_${cfg}_${n}gpu$tag.json 2> $O/r2y_${cfg}_${n}gpu$tag.err
  python - <<PY || tail -5 $O/r2y_${cfg}_${n}gpu$tag.err
import json
lines = [l for l in open('$O/r2y_${cfg}_${n}gpu$tag.json') if l.startswith('{')]
d = json.loads(lines[-1])
print('$cfg x$n $tag', '%.3e' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.3e' % d['e2e']['value'],
      'slab_bit_identical', d.get('slab_bit_identical'), len(lines), 'json line(s)')
PY
}
run 8 29514 c3 "" --slab-checks gpr_stiff
( CUDA_VISIBLE_DEVICES=0,1 run 2 29511 c4 "" --slab-checks none ) &
( CUDA_VISIBLE_DEVICES=2,3,4,5 run 4 29512 c4 "" --slab-checks none ) &
wait
run 8 29513 c4 "" --slab-checks gpr_stiff
du -sm $O
