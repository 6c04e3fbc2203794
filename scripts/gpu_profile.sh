#!/bin/bash
# bench line + ncu launch list + one full ncu capture of the dominant kernel.  Logs -> gpurun_out/.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/bench.err
tail -n 2 gpurun_out/bench.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit=$?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-kl_rows} -s ${NCU_SKIP:-3} -c 1 -o gpurun_out/prof_${NCU_TAG:-kl} -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"
ls -la gpurun_out | tail -n 20
