#!/bin/bash
# experiment: jump-kernel configurations
for c in 6 4 2; do
  echo "== streamed FDDM_JUMP_CTAS=$c"
  FDDM_JUMP_CTAS=$c timeout 120 python scripts/microbench.py jump 2>&1 | grep -E "categorical|Error|error" | cut -c1-200
done
echo "== bf16 V=32000"; timeout 120 python scripts/microbench.py jump --V 32000 --dtype bf16 2>&1 | grep -E "jump|Error|error" | cut -c1-200
echo "== f32 V=32000"; timeout 120 python scripts/microbench.py jump --V 32000 --dtype f32 2>&1 | grep -E "jump|Error|error" | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 200 -k "jump or sampler" 2>&1 | tail -3
