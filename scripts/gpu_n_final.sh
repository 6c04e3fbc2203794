#!/bin/bash
# Multi-GPU bench with default flags (what the driver runs) and, optionally, with the overlap forced on.
# usage: gpu_n_final.sh N [variants...]   variants: default overlap
N=${1:-2}; shift
VARS=${@:-default}
mkdir -p gpurun_out
for v in $VARS; do
  extra=""
  if [ $v = overlap ]; then extra="--collectives overlap"; fi
  t0=$(date +%s)
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 $extra > gpurun_out/final_${v}_n$N.json 2> gpurun_out/final_${v}_n$N.err
  echo "$v rc=$? after $(( $(date +%s) - t0 ))s: $(tail -n 1 gpurun_out/final_${v}_n$N.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['shard_check']['ok'], d['shard_check']['exchange'], d['e2e']['value'], d['config']['parallelism'][-60:])
except Exception as e: print('no line', e)")"
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/final_${v}_n$N.err | tail -n 3 | cut -c1-300
done
