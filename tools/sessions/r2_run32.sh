#!/bin/bash
# GPU session 32: pair wide kernel with the default-scope remote arrive
timeout 900 python -m pytest tests/test_gpu_gemm4.py -x -q -m gpu > gpurun_out/r32_t.log 2>&1; tail -3 gpurun_out/r32_t.log
timeout 300 python -m pytest tests/test_gpu_int8.py -x -q -m gpu 2>&1 | tail -2
for pm in 96 0; do
  echo "--- BNB_B200_GEMM4_WIDE_PAIR=$pm"
  BNB_B200_GEMM4_WIDE_PAIR=$pm timeout 300 python tools/kbench.py --only gemm4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if '_b128' in d['kernel'] or '_b256' in d['kernel']: print(d['kernel'], d['us'], d.get('TFLOPs'), d.get('bf16_frac'))"
done
