#!/bin/bash
# N-GPU A/B of the collective schedule: serial vs overlapped (several NCCL CTA counts).  usage: gpu_multi_ab.sh N
N=${1:-8}
mkdir -p gpurun_out
run() {  # tag, extra args / env
  tag=$1; shift
  t0=$(date +%s)
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu $ARGS > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
  echo "$tag exit=$? after $(( $(date +%s) - t0 ))s: $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_n${N}_$tag.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('shard_check',{}).get('ok'))
except Exception as e: print('unreadable', e)")"
}
ARGS="--collectives serial" run serial A=1
ARGS="--collectives overlap" run overlap8 NCCL_MAX_CTAS=8
ARGS="--collectives overlap" run overlap4 NCCL_MAX_CTAS=4
ARGS="--collectives overlap" run overlap16 NCCL_MAX_CTAS=16
