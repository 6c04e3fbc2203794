#!/bin/bash
# N-GPU A/B of the collective schedule: serial vs overlapped at several NCCL CTA counts.
# usage: gpu_multi_ab.sh N "serial 8 16 24"   (first round-2 result: profiles/r02h_n8_collective_schedule_ab.txt)
N=${1:-8}
mkdir -p gpurun_out
for v in ${2:-serial 16 24 32}; do
  if [ "$v" = serial ]; then tag=serial; args="--collectives serial"; envs="A=1"; else tag=overlap$v; args="--collectives overlap"; envs="NCCL_MAX_CTAS=$v"; fi
  t0=$(date +%s)
  env $envs timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu $args > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
  echo "$tag exit=$? after $(( $(date +%s) - t0 ))s: $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_$tag.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('shard_check',{}).get('ok'))
except Exception as e: print('unreadable', e)")"
done
