#!/bin/bash
# Final multi-GPU bench with default flags (what the driver runs), plus one variant of the reserved-SM count.
N=${1:-8}
mkdir -p gpurun_out
for v in default ctas4; do
  if [ $v = ctas4 ]; then export FDDM_XGPU_CTAS=4; fi
  t0=$(date +%s)
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_${v}_n$N.json 2> gpurun_out/final_${v}_n$N.err
  echo "$v rc=$? after $(( $(date +%s) - t0 ))s: $(tail -n 1 gpurun_out/final_${v}_n$N.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['shard_check']['ok'], d['shard_check']['exchange'], d['e2e']['value'])
except Exception as e: print('no line', e)")"
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/final_${v}_n$N.err | tail -n 3
done
