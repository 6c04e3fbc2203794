#!/bin/bash
# First-contact GPU check: smoke, then the gpu-marked parity tests in two processes (the tcgen05 L_fd tests
# are isolated so that a fault there cannot hide the row-kernel results).  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "not lfd" > gpurun_out/test_rows.log 2>&1
echo "rows exit=$?" >> gpurun_out/test_rows.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "lfd" > gpurun_out/test_lfd.log 2>&1
echo "lfd exit=$?" >> gpurun_out/test_lfd.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 600 > gpurun_out/test_full.log 2>&1
echo "full exit=$?" >> gpurun_out/test_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?" >> gpurun_out/smoke.log
for f in test_rows test_lfd test_full smoke; do tail -n 4 gpurun_out/$f.log; done
