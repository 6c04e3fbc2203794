#!/bin/bash
# experiment: jump-kernel configurations (consumer threads x entries per thread x CTAs/SM [x stages])
for cfg in 256x32x3 128x64x3 128x64x4x1 256x32x3x1 128x64x3x3; do
  echo "== FDDM_JUMP_CFG=$cfg"
  FDDM_JUMP_CFG=$cfg timeout 120 python scripts/microbench.py jump 2>&1 | grep -E "categorical|Error|error" | cut -c1-200
done
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 200 -k "jump or sampler" 2>&1 | tail -3
