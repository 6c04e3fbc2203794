#!/bin/bash
# Bench lines for every BASELINE config on one GPU -> gpurun_out/bench_<tag>.json, then the ncu launch list of the
# default bench command (eager launches, so every kernel is visible).
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "$tag exit=$?"; }
run c5 
run c5shard --workload c5shard --no-cpu
run c2 --workload c2 --no-cpu
run c2_bf16 --workload c2 --dtype bf16 --no-cpu
run c4_f32 --workload c4 --no-cpu
run c4_bf16 --workload c4 --dtype bf16 --no-cpu
run c3_exact_cat --workload c3 --sampling-mode exact
run c3_exact_greedy --workload c3 --sampling-mode exact --greedy --no-cpu
run c3_fast_cat --workload c3 --sampling-mode fast --no-cpu
run c3_fast_greedy --workload c3 --sampling-mode fast --greedy --no-cpu
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_c*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f.split("bench_")[1][:-5].ljust(16), d["value"], "G", d["ms_per_step"], "ms  e2e", d["e2e"]["value"], " step_frac", r["step_frac"],
              " dom", (r["kernel"] or "")[:24], r["frac"], " eager", (d.get("eager_b200") or {}).get("value"), " cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_c5.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit=$?"
