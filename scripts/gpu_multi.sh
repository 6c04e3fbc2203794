#!/bin/bash
# N-GPU bench (torchrun, NCCL).  usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  scripts/check_multi_gpu.py 2>&1 | grep -E "multi-gpu check|Error|error" | tee gpurun_out/check_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit=$?"
tail -n 1 gpurun_out/bench_n$N.json | cut -c1-400
tail -n 5 gpurun_out/bench_n$N.err
