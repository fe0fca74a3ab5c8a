#!/bin/bash
# GPU session 28: wide kernel, four dequant groups (NB <= 192) vs three, same box
timeout 900 python -m pytest tests/test_gpu_gemm4.py -q -m gpu > gpurun_out/r28_t.log 2>&1; tail -3 gpurun_out/r28_t.log
timeout 300 python tools/gemm4_stress.py 40 2>&1 | tail -6
for g in 4 3; do
  BNB_B200_GEMM4_WIDE_G=$g BNB_B200_GEMM4_SMALL=0 timeout 300 python tools/kbench.py --only gemm4 > gpurun_out/r28_kbench_gemm4_g$g.jsonl 2>&1
  echo "--- wide kernel at every batch, G=$g"
  python - <<PY
import json
for l in open('gpurun_out/r28_kbench_gemm4_g$g.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kernel'], d['us'], d.get('TFLOPs'), d.get('bf16_frac'), d.get('speedup_vs_composition'))
PY
done
