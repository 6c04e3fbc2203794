#!/bin/bash
# A/B of the L_fd exchange on N GPUs: NCCL vs the library's own symmetric-memory all-reduce kernels.
# usage: gpu_xgpu_ab.sh N [workload]
N=${1:-2}; W=${2:-c5}
mkdir -p gpurun_out
run() { # name, extra args
  name=$1; shift
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $W --steps 20 --warmup 5 "$@" > gpurun_out/ab_${name}_n$N.json 2> gpurun_out/ab_${name}_n$N.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${name}_n$N.json").read().strip().splitlines()[-1])
    print(d.get("value"), d.get("ms_per_step"), (d.get("shard_check") or {}).get("ok"), (d.get("shard_check") or {}).get("exchange"), d.get("error", ""))
except Exception as e:
    print("no line:", e)
PY
)"
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/ab_${name}_n$N.err | tail -n 3
}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  scripts/nvls_probe.py > gpurun_out/xgpu_probe_n$N.json 2> gpurun_out/xgpu_probe_n$N.err; echo "probe rc=$?"
if [ -z "$SKIP_CHECK" ]; then
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    scripts/check_multi_gpu.py > gpurun_out/check_n$N.log 2>&1
  echo "check rc=$?"; grep -E "multi-gpu check" gpurun_out/check_n$N.log | cut -c1-200 | tee gpurun_out/check_n$N.txt
fi
run nccl_auto --exchange nccl
if [ -z "$SKIP_SERIAL" ]; then run p2p_serial --exchange p2p --collectives serial; fi
run p2p_overlap --exchange p2p --collectives overlap
if [ -z "$SKIP_NVLS" ]; then
  run nvls_overlap --exchange nvls --collectives overlap
  NCCL_MAX_CTAS=8 run nvls_overlap_8ctas --exchange nvls --collectives overlap
fi
