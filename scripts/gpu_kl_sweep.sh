#!/bin/bash
# experiment: KL kernel variants
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 200 -k "kl" 2>&1 | tail -4
echo "== register-resident (old)"; FDDM_KL_REG=1 timeout 120 python scripts/microbench.py kl 2>&1 | grep -E "kl|rror" | cut -c1-200
for c in 6 5 4 3; do
  echo "== streamed CTAS=$c"; FDDM_KL_CTAS=$c timeout 120 python scripts/microbench.py kl 2>&1 | grep -E "kl|rror" | cut -c1-200
done
echo "== streamed CTAS=3 STAGES=2"; FDDM_KL_CTAS=3 FDDM_KL_STAGES=2 timeout 120 python scripts/microbench.py kl 2>&1 | grep -E "kl|rror" | cut -c1-200
echo "== V=32000 f32 / bf16"; timeout 120 python scripts/microbench.py kl --V 32000 2>&1 | grep -E "kl|rror" | cut -c1-200; timeout 120 python scripts/microbench.py kl --V 32000 --dtype bf16 2>&1 | grep -E "kl|rror" | cut -c1-200
echo "== V=8000 bf16"; timeout 120 python scripts/microbench.py kl --dtype bf16 2>&1 | grep -E "kl|rror" | cut -c1-200
