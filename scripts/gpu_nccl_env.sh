#!/bin/bash
# all-reduce latency at the step's message sizes under a few NCCL settings (8 GPUs)
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/nccl_probe.py 2>&1 | grep allreduce; }
{ run X=1; run NCCL_ALGO=NVLS; run NCCL_ALGO=Tree; run NCCL_PROTO=LL128; run NCCL_MAX_NCHANNELS=32 NCCL_MIN_NCHANNELS=32; } > gpurun_out/nccl_env_n8.txt 2>&1
cat gpurun_out/nccl_env_n8.txt
