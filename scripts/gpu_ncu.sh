#!/bin/bash
# ncu --set full captures of the hot kernels (one launch each) on the c5shard shapes, after the same command
# has exited 0 without ncu.  Reports -> gpurun_out/prof_*.ncu-rep (read here with ncu -i).
mkdir -p gpurun_out
CMD="python scripts/microbench.py jump kl lfd --iters 3"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
for k in ${NCU_KERNELS:-kl_rows_ring jump_rows_streamed umma_bwd_kernel}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_r2_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k exit=$?"
done
ls -la gpurun_out/*.ncu-rep
