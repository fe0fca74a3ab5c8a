#!/bin/bash
# GPU session 31: wide 4-bit GEMM with CTA pairs (batch >= 96) -- tests, cold-cache stress, same-box A/B
timeout 900 python -m pytest tests/test_gpu_gemm4.py -x -q -m gpu > gpurun_out/r31_t.log 2>&1; tail -4 gpurun_out/r31_t.log
timeout 300 python tools/gemm4_stress.py 40 2>&1 | tail -6
for pm in 96 0 48; do
  echo "--- BNB_B200_GEMM4_WIDE_PAIR=$pm"
  BNB_B200_GEMM4_WIDE_PAIR=$pm timeout 300 python tools/kbench.py --only gemm4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if '_b16' in d['kernel'] or '_b32' in d['kernel']: continue
    print(d['kernel'], d['us'], d.get('TFLOPs'), d.get('bf16_frac'), d.get('speedup_vs_composition'))"
done
