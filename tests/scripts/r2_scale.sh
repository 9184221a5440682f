# usage: bash tests/scripts/r2_scale.sh N   — copy ceiling + bench (+ reference arm) at N GPUs of one box
N=$1
P=29500
if [ "$N" = "1" ]; then
  python tests/gpu_copyceiling.py > gpurun_out/r2_ceiling_n$N.json 2> gpurun_out/r2_ceiling_n$N.err
  python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tests/gpu_copyceiling.py > gpurun_out/r2_ceiling_n$N.json 2> gpurun_out/r2_ceiling_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
fi
cat gpurun_out/r2_ceiling_n$N.json; tail -2 gpurun_out/r2_ceiling_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e_pageable"]["value"]), "compress", d["compress_e2e"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/r2_bench_n$N.err
nproc; free -g | head -2
