# N-GPU e2e variants: pipeline depth x host wait mode (bench without the config / compress / cpu legs)
N=$1
P=29600
i=0
for cfg in "6 0" "4 0" "4 1" "3 0" "8 0"; do
  set -- $cfg
  i=$((i+1))
  BWTC_WAIT_MODE=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+i)) bench.py --gpus $N --steps 8 --warmup 3 --depth $1 --no-configs --no-compress --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=$N depth $1 wait_mode $2: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pageable', round(d['e2e_pageable']['value']))"
done
