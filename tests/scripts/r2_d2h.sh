for v in "0 32" "1 16" "1 32" "1 64" "1 128"; do
  set -- $v
  BWTC_D2H_KERNEL=$1 BWTC_D2H_CTAS=$2 python bench.py --steps 8 --warmup 3 --no-configs --no-compress --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('D2H_KERNEL=$1 CTAS=$2: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pageable', round(d['e2e_pageable']['value']))"
done
python -m pytest tests/test_gpu_host_paths.py tests/test_gpu_parity.py -q -x -k "not maximum_block and not lazy" 2>&1 | tail -3
