#!/bin/bash
# First GPU call of the next round: re-validate the shipped path, then try the experimental persistent
# tcgen05 contraction (FDDM_UMMA_PERSISTENT=1) on the L_fd tests and the bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
bash scripts/gpu_check.sh
timeout 600 python bench.py --no-cpu > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit=$?"
FDDM_UMMA_PERSISTENT=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --timeout 300 -k "lfd" > gpurun_out/test_lfd_persistent.log 2>&1
echo "lfd persistent exit=$?"; tail -n 3 gpurun_out/test_lfd_persistent.log
FDDM_UMMA_PERSISTENT=1 timeout 600 python bench.py --no-cpu > gpurun_out/bench_persistent.json 2> gpurun_out/bench_persistent.err
echo "bench persistent exit=$?"
python - <<'PY'
import json
for tag in ("default", "persistent"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{tag}.json").read().strip().splitlines()[-1])
        print(tag, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["cuda_graph"])
    except Exception as e:
        print(tag, "unreadable:", e)
PY
