for d in 4 6 8 12; do
  python bench.py --steps 8 --warmup 3 --depth $d --no-configs --no-compress --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=1 depth $d: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pageable', round(d['e2e_pageable']['value']))"
done
