#!/bin/bash
# N-GPU correctness check + bench (torchrun, NCCL), every step under its own short timeout so that a hang
# costs minutes, not the round's budget.  usage: gpu_multi.sh N [extra bench args]
N=${1:-2}; shift
mkdir -p gpurun_out
t0=$(date +%s)
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  scripts/check_multi_gpu.py > gpurun_out/check_n$N.log 2>&1
rc=$?; grep -E "multi-gpu check" gpurun_out/check_n$N.log | tee gpurun_out/check_n$N.txt
echo "check N=$N exit=$rc after $(( $(date +%s) - t0 ))s"
if [ $rc -ne 0 ]; then tail -n 20 gpurun_out/check_n$N.log; echo "check FAILED: not benchmarking"; exit 1; fi
t0=$(date +%s)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
rc=$?
echo "bench N=$N exit=$rc after $(( $(date +%s) - t0 ))s"
tail -n 1 gpurun_out/bench_n$N.json | cut -c1-600
grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/bench_n$N.err | tail -n 8
exit $rc
