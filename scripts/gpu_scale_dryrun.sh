#!/bin/bash
# What the driver runs for SCALE, for the given GPU counts (default flags): reference arm first, then the B200 arm.
# usage: gpu_scale_dryrun.sh "2 4"
mkdir -p gpurun_out
for N in ${1:-2}; do
  for impl in reference b200; do
    t0=$(date +%s)
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --impl $impl --gpus 1 --steps 20 --warmup 5 > gpurun_out/scale_${impl}_n$N.json 2> gpurun_out/scale_${impl}_n$N.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --impl $impl --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_${impl}_n$N.json 2> gpurun_out/scale_${impl}_n$N.err
    fi
    echo "N=$N impl=$impl exit=$? after $(( $(date +%s) - t0 ))s: $(tail -n 1 gpurun_out/scale_${impl}_n$N.json | cut -c1-260)"
  done
done
