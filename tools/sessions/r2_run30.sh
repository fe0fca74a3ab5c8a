#!/bin/bash
# GPU session 30: int8 GEMM with CTA pairs (tcgen05 cta_group::2) -- parity tests, then same-box A/B against one CTA per MMA
timeout 600 python -m pytest tests/test_gpu_int8.py -x -q -m gpu > gpurun_out/r30_t.log 2>&1; tail -5 gpurun_out/r30_t.log
for cg in 2 1; do
  echo "--- BNB_B200_IGEMM_CG=$cg"
  BNB_B200_IGEMM_CG=$cg timeout 300 python tools/kbench.py --only int8 2>&1 | grep -E "igemm|Linear8|rror" | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['us'], d.get('TOPS'))"
done
